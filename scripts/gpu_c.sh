#!/bin/bash
set -u
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for w in ransac loop dense; do
  python bench.py --workload $w > gpurun_out/bench_${w}_1gpu.json 2> gpurun_out/bench_${w}.err; echo "$w rc=$?"; tail -3 gpurun_out/bench_${w}.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${w}_1gpu.json').read())
    print('$w', 'value', d['value'], d['unit'], 'ms', round(d['ms_per_step'],2), 'e2e', d['e2e']['value'], 'roof', round(d['roofline']['frac'],3), 'cpu', d['cpu_baseline']['value'], d['parity'])
except Exception as e: print('$w parse failed', e)
PY
done
