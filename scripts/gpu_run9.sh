set -x
nvidia-smi -L
timeout 1300 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_all.log 2>&1
tail -25 gpurun_out/r02_pytest_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -4 gpurun_out/r02_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; tail -c 6000 gpurun_out/r02_bench_a.json; tail -5 gpurun_out/r02_bench_a.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; tail -c 2000 gpurun_out/r02_bench_ref.json; tail -5 gpurun_out/r02_bench_ref.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-kernel-table > gpurun_out/r02_ncu_bench.log 2>&1; tail -2 gpurun_out/r02_ncu_bench.log
LONG=1 timeout 300 python scripts/bench_attn.py > gpurun_out/r02_attn_ab.log 2>&1; cat gpurun_out/r02_attn_ab.log
TAG=default timeout 120 python scripts/bench_attn_vit.py > gpurun_out/r02_attn_vit.log 2>&1; cat gpurun_out/r02_attn_vit.log
