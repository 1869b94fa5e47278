S=$(date +%s)
timeout 100 python -c "import torch; print('torch ok', torch.cuda.device_count())" || exit 7
[ $(( $(date +%s) - S )) -gt 60 ] && { echo "slow box: abort"; exit 7; }
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_n8_ep.json 2> gpurun_out/r02_bench_n8_ep.err; tail -c 500 gpurun_out/r02_bench_n8_ep.json; tail -4 gpurun_out/r02_bench_n8_ep.err
