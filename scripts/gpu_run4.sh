set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "attention" > gpurun_out/r02_attn_tests.log 2>&1; tail -15 gpurun_out/r02_attn_tests.log
ARIA_ATTN_W=128 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "attention" > gpurun_out/r02_attn_tests_w128.log 2>&1; tail -5 gpurun_out/r02_attn_tests_w128.log
timeout 600 python -m pytest tests/test_gpu_parity_full.py -q -x -s -k "attention" > gpurun_out/r02_attn_tests_full.log 2>&1; tail -15 gpurun_out/r02_attn_tests_full.log
LONG=1 timeout 300 python scripts/bench_attn.py > gpurun_out/r02_attn_ab.log 2>&1
ARIA_ATTN_W=128 timeout 300 python scripts/bench_attn.py >> gpurun_out/r02_attn_ab.log 2>&1
ARIA_ATTN_PERSIST=0 timeout 300 python scripts/bench_attn.py >> gpurun_out/r02_attn_ab.log 2>&1
cat gpurun_out/r02_attn_ab.log
