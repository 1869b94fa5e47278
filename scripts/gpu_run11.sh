nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv -lms 500 > gpurun_out/r02_abl_clocks.csv &
SMI=$!
TAG="full kernel" python scripts/bench_attn_vit.py > gpurun_out/r02_attn_ablation.log 2>&1
for v in 1 2 4 8 16 3 19 31; do TAG="ablate=$v" ARIA_B200_LIB=$PWD/aria_b200/build/libaria_abl$v.so timeout 120 python scripts/bench_attn_vit.py >> gpurun_out/r02_attn_ablation.log 2>&1; done
for v in poly4 poly8; do TAG="$v" ARIA_B200_LIB=$PWD/aria_b200/build/libaria_$v.so timeout 120 python scripts/bench_attn_vit.py >> gpurun_out/r02_attn_ablation.log 2>&1; done
kill $SMI
TAG=full ARIA_B200_LIB=$PWD/aria_b200/build/libaria_trace.so python scripts/trace_attn.py >> gpurun_out/r02_attn_ablation.log 2>&1
cat gpurun_out/r02_attn_ablation.log
sort -t, -k1 -n gpurun_out/r02_abl_clocks.csv | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd3 -s 1 -c 1 -o gpurun_out/r02_attn_vit python scripts/prof_kernels.py attn > gpurun_out/r02_ncu_attn.log 2>&1; tail -3 gpurun_out/r02_ncu_attn.log
