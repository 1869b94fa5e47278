set -x
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv -lms 500 > gpurun_out/r02_abl_clocks.csv &
SMI=$!
TAG="full kernel (poly4, seq)" python scripts/bench_attn_vit.py > gpurun_out/r02_attn_ablation.log 2>&1
for v in 1 2 3 4 8 12 16 19 31; do TAG="ablate=$v" ARIA_B200_LIB=$PWD/aria_b200/build/libaria_abl$v.so timeout 120 python scripts/bench_attn_vit.py >> gpurun_out/r02_attn_ablation.log 2>&1; done
kill $SMI
cat gpurun_out/r02_attn_ablation.log
sort -t, -k1 -n gpurun_out/r02_abl_clocks.csv | tail -3
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "permutation or attention or moe_layer" 2>&1 | tail -4
