set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py -q -x -k "attention" > gpurun_out/r02_attn_tests.log 2>&1; tail -5 gpurun_out/r02_attn_tests.log
LONG=1 timeout 300 python scripts/bench_attn.py > gpurun_out/r02_attn_ab2.log 2>&1
for v in poly0 poly6 poly8; do ARIA_B200_LIB=$PWD/aria_b200/build/libaria_$v.so timeout 300 python scripts/bench_attn.py 2>&1 | sed "s/^/[$v] /" >> gpurun_out/r02_attn_ab2.log; done
cat gpurun_out/r02_attn_ab2.log
