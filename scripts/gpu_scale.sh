#!/bin/bash
# Bench workloads at N GPUs: gpurun --gpus N --timeout 600 -- 'bash scripts/gpu_scale.sh N "sequence loop"'
set -u
N=${1:-2}
WL=${2:-"sequence ransac loop dense"}
for w in $WL; do
  if [ "$N" = "1" ]; then
    timeout 300 python bench.py --workload $w --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${w}_${N}gpu.json 2> gpurun_out/scale_${w}_${N}gpu.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --workload $w --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${w}_${N}gpu.json 2> gpurun_out/scale_${w}_${N}gpu.err
  fi
  echo "$w rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/scale_${w}_${N}gpu.json').read().strip().splitlines()[-1])
    print('$w', 'N', d['n_gpus'], 'value', d['value'], d['unit'], 'ms', round(d['ms_per_step'],2), 'e2e', d['e2e']['value'], d['scaling'])
except Exception as e: print('$w parse failed', e)
PY
done
