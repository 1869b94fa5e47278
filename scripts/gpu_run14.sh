TAG=w80 timeout 100 python scripts/bench_attn_vit.py > gpurun_out/r02_attn_ab6.log 2>&1
TAG=w128 ARIA_ATTN_W=128 timeout 100 python scripts/bench_attn_vit.py >> gpurun_out/r02_attn_ab6.log 2>&1
TAG=w80-nopersist ARIA_ATTN_PERSIST=0 timeout 100 python scripts/bench_attn_vit.py >> gpurun_out/r02_attn_ab6.log 2>&1
cat gpurun_out/r02_attn_ab6.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd3 -s 1 -c 1 -o gpurun_out/r02_attn_vit72 python scripts/prof_kernels.py attn72 > gpurun_out/r02_ncu_attn72.log 2>&1; tail -3 gpurun_out/r02_ncu_attn72.log
