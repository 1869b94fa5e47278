timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py tests/test_install_vit.py tests/test_hf_attention.py -q -x -k "attention or vit or hf" 2>&1 | tail -3
LONG=1 python scripts/bench_attn.py > gpurun_out/r02_attn_ab4.log 2>&1
TAG=roles python scripts/bench_attn_vit.py >> gpurun_out/r02_attn_ab4.log 2>&1
TAG=full ARIA_B200_LIB=$PWD/aria_b200/build/libaria_trace.so python scripts/trace_attn.py >> gpurun_out/r02_attn_ab4.log 2>&1
cat gpurun_out/r02_attn_ab4.log
