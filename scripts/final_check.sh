#!/bin/bash
# Round-end style verification on the GPU box: parity tests, smoke, reference arm, bench (writes gpurun_out/).
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -2 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1 | tee gpurun_out/smoke.log
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.json | cut -c1-250
python bench.py 2>&1 | tail -1 > gpurun_out/bench_final.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_final.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")})
print(d["e2e"])
print({k: d["roofline"][k] for k in ("achieved", "frac", "avg_launch_ms", "share_of_step", "traffic")})
print(d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
