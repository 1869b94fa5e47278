S=$(date +%s)
timeout 100 python -c "import torch; print('torch ok', torch.cuda.device_count())" || exit 7
[ $(( $(date +%s) - S )) -gt 60 ] && { echo "slow box: abort"; exit 7; }
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pair_mode or grouped_gemm or moe_layer" 2>&1 | tail -4
timeout 100 python scripts/prof_kernels.py ep_fc1_expert_major 2>&1 | tail -1
ARIA_GEMM_PAIR=0 timeout 100 python scripts/prof_kernels.py ep_fc1_expert_major 2>&1 | tail -1
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:gemm -s 3 -c 1 python scripts/prof_kernels.py ep_fc1_expert_major 2>&1 | grep -E "gemm|dram__bytes|gpu__time" | tail -3
