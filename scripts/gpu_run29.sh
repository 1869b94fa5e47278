S=$(date +%s)
timeout 100 python -c "import torch; print('torch ok')" || exit 7
[ $(( $(date +%s) - S )) -gt 60 ] && { echo "slow box: abort"; exit 7; }
TAG=heads2 timeout 120 python scripts/bench_gemm_shapes.py 2>&1 | grep -E "qkv|o_proj"
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err; tail -c 200 gpurun_out/r02_bench_e.json; tail -2 gpurun_out/r02_bench_e.err
