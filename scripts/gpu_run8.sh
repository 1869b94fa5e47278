timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py -q -x -k "attention" 2>&1 | tail -3
LONG=1 python scripts/bench_attn.py > gpurun_out/r02_attn_ab3.log 2>&1
for v in poly2 poly4 seq; do TAG=$v ARIA_B200_LIB=$PWD/aria_b200/build/libaria_$v.so python scripts/bench_attn_vit.py >> gpurun_out/r02_attn_ab3.log 2>&1; done
TAG=full ARIA_B200_LIB=$PWD/aria_b200/build/libaria_trace.so python scripts/trace_attn.py >> gpurun_out/r02_attn_ab3.log 2>&1
cat gpurun_out/r02_attn_ab3.log
