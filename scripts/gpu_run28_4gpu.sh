S=$(date +%s)
timeout 100 python -c "import torch; print('torch ok', torch.cuda.device_count())" || exit 7
[ $(( $(date +%s) - S )) -gt 60 ] && { echo "slow box: abort"; exit 7; }
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 5 --warmup 3 --no-kernel-table > gpurun_out/r02_bench_n4_ep.json 2> gpurun_out/r02_bench_n4_ep.err; tail -c 600 gpurun_out/r02_bench_n4_ep.json; tail -4 gpurun_out/r02_bench_n4_ep.err
