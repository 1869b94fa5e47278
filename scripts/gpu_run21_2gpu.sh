S=$(date +%s)
timeout 150 python -c "import torch; print('torch ok', torch.cuda.device_count())" || exit 7
[ $(( $(date +%s) - S )) -gt 100 ] && { echo "slow box: abort"; exit 7; }
timeout 600 python -m pytest tests/test_gpu_ep.py -m gpu -q -x 2>&1 | tail -4
P=29521
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2_ep_b.json 2> gpurun_out/r02_bench_n2_ep_b.err; tail -c 700 gpurun_out/r02_bench_n2_ep_b.json; tail -4 gpurun_out/r02_bench_n2_ep_b.err
