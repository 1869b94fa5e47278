nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_ep.py -m gpu -q -x 2>&1 | tail -6
timeout 300 python -m pytest tests/test_gpu_dropin.py -m gpu -q -x -k "two_devices" 2>&1 | tail -3
P=29511
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2_ep.json 2> gpurun_out/r02_bench_n2_ep.err; tail -c 1200 gpurun_out/r02_bench_n2_ep.json; tail -4 gpurun_out/r02_bench_n2_ep.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus 2 --steps 5 --warmup 3 --workload cfg5 > gpurun_out/r02_bench_n2_cfg5.json 2> gpurun_out/r02_bench_n2_cfg5.err; tail -c 1500 gpurun_out/r02_bench_n2_cfg5.json; tail -4 gpurun_out/r02_bench_n2_cfg5.err
