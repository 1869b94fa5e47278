S=$(date +%s)
timeout 100 python -c "import torch; print('torch ok', torch.cuda.device_count())" || exit 7
[ $(( $(date +%s) - S )) -gt 60 ] && { echo "slow box: abort"; exit 7; }
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1; tail -3 gpurun_out/r02_smoke_final.log
timeout 500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_final.log 2>&1; tail -3 gpurun_out/r02_pytest_final.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r02_clocks_final.csv &
SMI=$!
timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -c 400 gpurun_out/r02_bench_final.json; tail -3 gpurun_out/r02_bench_final.err
kill $SMI
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_final.json 2> gpurun_out/r02_bench_ref_final.err; tail -c 300 gpurun_out/r02_bench_ref_final.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 1900 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-kernel-table > gpurun_out/r02_ncu_bench_final.log 2>&1; tail -1 gpurun_out/r02_ncu_bench_final.log | cut -c1-200
timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_fwd3 -s 1 -c 1 -o gpurun_out/r02_attn_vit72_final python scripts/prof_kernels.py attn72 > gpurun_out/r02_ncu_attn72_final.log 2>&1; tail -1 gpurun_out/r02_ncu_attn72_final.log
timeout 200 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline --no-kernel-table > gpurun_out/r02_bench_cfg3.json 2> gpurun_out/r02_bench_cfg3.err; tail -c 300 gpurun_out/r02_bench_cfg3.json; tail -2 gpurun_out/r02_bench_cfg3.err
timeout 300 python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-table > gpurun_out/r02_bench_cfg4.json 2> gpurun_out/r02_bench_cfg4.err; tail -c 300 gpurun_out/r02_bench_cfg4.json; tail -2 gpurun_out/r02_bench_cfg4.err
