#!/bin/bash
set -u
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1; echo "kernel_bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/kernels.json'))
for k,v in d.items():
    if isinstance(v,dict): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a not in ('note','bound')})
    else: print(k,v)
PY
for c in 576; do python bench.py --no-cpu-baseline --chunk-frames $c 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('chunk',$c,'value',round(d['value']),'e2e',round(d['e2e']['value']),'e2e_ms',round(d['e2e']['ms_per_step'],1),'ms',round(d['ms_per_step'],1), 'launch_ms', round(d['roofline']['launch_ms'],1))"; done
