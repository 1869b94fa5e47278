S=$(date +%s)
timeout 100 python -c "import torch; print('torch ok')" || exit 7
[ $(( $(date +%s) - S )) -gt 60 ] && { echo "slow box: abort"; exit 7; }
TAG=default timeout 120 python scripts/bench_gemm_shapes.py 2>&1 | grep "LM "
TAG=mid ARIA_GEMM_MID=1 timeout 120 python scripts/bench_gemm_shapes.py 2>&1 | grep "LM "
TAG=1cta ARIA_GEMM_CTAS=1 timeout 120 python scripts/bench_gemm_shapes.py 2>&1 | grep "ViT "
