S=$(date +%s)
timeout 100 python -c "import torch; print('torch ok')" || exit 7
[ $(( $(date +%s) - S )) -gt 60 ] && { echo "slow box: abort"; exit 7; }
for m in dense dense_oproj dense_lm; do
timeout 150 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.per_cycle_active,dram__bytes_read.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,l1tex__m_xbar2l1tex_read_bytes.sum --clock-control none -k regex:gemm -s 2 -c 1 python scripts/prof_kernels.py $m 2>&1 | grep -E "gemm|gpu__time|tensor|issue_active|dram__|long_score|xbar" | sed "s/^/[$m] /"
done
