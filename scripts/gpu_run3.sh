set -x
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest_all.log 2>&1
tail -25 gpurun_out/r02_pytest_all.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; tail -c 3000 gpurun_out/r02_bench_a.json; tail -5 gpurun_out/r02_bench_a.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; tail -c 1500 gpurun_out/r02_bench_ref.json; tail -5 gpurun_out/r02_bench_ref.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd2 -s 1 -c 1 -o gpurun_out/r02_attn_vit python scripts/prof_kernels.py attn > gpurun_out/r02_ncu_attn.log 2>&1; tail -3 gpurun_out/r02_ncu_attn.log
