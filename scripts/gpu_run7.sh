TAG=full ARIA_B200_LIB=$PWD/aria_b200/build/libaria_trace.so python scripts/trace_attn.py > gpurun_out/r02_attn_trace.log 2>&1
TAG=ablate31 ARIA_B200_LIB=$PWD/aria_b200/build/libaria_trace31.so python scripts/trace_attn.py >> gpurun_out/r02_attn_trace.log 2>&1
cat gpurun_out/r02_attn_trace.log
