S=$(date +%s)
timeout 150 python -c "import torch; print('torch ok', torch.cuda.device_count())" || exit 7
[ $(( $(date +%s) - S )) -gt 100 ] && { echo "slow box: abort"; exit 7; }
timeout 100 python scripts/prof_kernels.py ep_fc1_expert_major 2>&1 | tail -1
timeout 100 python scripts/prof_kernels.py ep_fc1_source_major 2>&1 | tail -1
TAG=new2 timeout 150 python scripts/bench_gemm_shapes.py > gpurun_out/r02_gemm_shapes2.log 2>&1; cat gpurun_out/r02_gemm_shapes2.log | tail -12
timeout 500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err; tail -c 300 gpurun_out/r02_bench_d.json; tail -3 gpurun_out/r02_bench_d.err
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:gemm_kernel -s 3 -c 1 python scripts/prof_kernels.py ep_fc1_expert_major 2>&1 | grep -E "dram__bytes|gpu__time" | tail -2
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:gemm_kernel -s 3 -c 1 python scripts/prof_kernels.py ep_fc1_source_major 2>&1 | grep -E "dram__bytes|gpu__time" | tail -2
