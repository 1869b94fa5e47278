S=$(date +%s)
timeout 100 python -c "import torch; print('torch ok')" || exit 7
[ $(( $(date +%s) - S )) -gt 60 ] && { echo "slow box: abort"; exit 7; }
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_final2.json 2> gpurun_out/r02_bench_final2.err; tail -c 200 gpurun_out/r02_bench_final2.json; tail -2 gpurun_out/r02_bench_final2.err
python -c "
import json
l=json.loads(open('gpurun_out/r02_bench_final2.json').read().strip().splitlines()[-1])
print('RESULT', l['value'], l['ms_per_step'], l['e2e']['value'], l['e2e']['ms_per_step'], l['clocks'])
for k in l['kernels']: print('K', k['kernel'], round(k['ms_per_step'],3), round(k.get('avg_launch_us',0),1), round(k['share'],3), round(k.get('frac',0) or 0,3))
"
