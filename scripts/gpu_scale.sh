#!/bin/bash
# Default bench (sequence + loop / dense sub-records) at the box's GPU count.  Run as:
#   gpurun --gpus N --timeout 900 -- 'bash scripts/gpu_scale.sh N'
set -u
N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_scale_${N}gpu.json 2> gpurun_out/r02_scale_${N}gpu.err; echo "bench N=$N rc=$?"
python - <<PY
import json
txt=open('gpurun_out/r02_scale_${N}gpu.json').read()
d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print('N',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e'].get('pcie'))
for k,v in d['extra'].items():
    print(k,v['value'],'ms',v['ms_per_step'],'e2e',v['e2e']['value'],json.dumps(v['parity'])[:400])
PY
