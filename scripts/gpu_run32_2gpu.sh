S=$(date +%s)
timeout 100 python -c "import torch; print('torch ok', torch.cuda.device_count())" || exit 7
[ $(( $(date +%s) - S )) -gt 60 ] && { echo "slow box: abort"; exit 7; }
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 5 --warmup 3 --no-kernel-table > gpurun_out/r02_bench_n2_ep_final.json 2> gpurun_out/r02_bench_n2_ep_final.err; tail -c 400 gpurun_out/r02_bench_n2_ep_final.json; tail -2 gpurun_out/r02_bench_n2_ep_final.err
timeout 140 python -m pytest tests/test_gpu_ep.py -m gpu -q -x -k "fused or backward" 2>&1 | tail -3
