timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
LONG=1 timeout 300 python scripts/bench_attn.py > gpurun_out/r02_attn_ab7.log 2>&1; cat gpurun_out/r02_attn_ab7.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; tail -c 1500 gpurun_out/r02_bench_b.json; tail -5 gpurun_out/r02_bench_b.err
