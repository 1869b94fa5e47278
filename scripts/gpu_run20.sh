S=$(date +%s)
timeout 150 python -c "import torch, time; t=time.time(); x=torch.randn(1024,1024,device='cuda'); torch.cuda.synchronize(); print('torch ok', time.time()-t)" || { echo "slow or broken box: abort"; exit 7; }
echo "import took $(( $(date +%s) - S )) s"
[ $(( $(date +%s) - S )) -gt 100 ] && { echo "slow box: abort"; exit 7; }
TAG=new timeout 150 python scripts/bench_gemm_shapes.py > gpurun_out/r02_gemm_shapes.log 2>&1; cat gpurun_out/r02_gemm_shapes.log | tail -12
timeout 500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; tail -c 600 gpurun_out/r02_bench_c.json; tail -3 gpurun_out/r02_bench_c.err
