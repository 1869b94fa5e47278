TAG=new timeout 300 python scripts/bench_gemm_shapes.py > gpurun_out/r02_gemm_shapes.log 2>&1
cat gpurun_out/r02_gemm_shapes.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; tail -c 600 gpurun_out/r02_bench_c.json; tail -3 gpurun_out/r02_bench_c.err
