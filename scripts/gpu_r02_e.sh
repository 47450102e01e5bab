#!/bin/bash
# Round 2, visit E (2 GPUs): GPU suite, createdb workload, default bench at N=1 and N=2 (sub-records with N>1 parity).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02e_pytest.log
python bench.py --workload createdb --steps 2 > gpurun_out/r02e_createdb.json 2> gpurun_out/r02e_createdb.err; echo "createdb rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02e_bench_2gpu.json 2> gpurun_out/r02e_bench_2gpu.err; echo "bench 2gpu rc=$?"
cat gpurun_out/r02e_createdb.json | cut -c1-1800
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02e_bench_2gpu.json'))
print('N=2 value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'])
for k,v in d['extra'].items():
    print(k,v['value'],'e2e',v['e2e']['value'],json.dumps(v['parity'])[:600])
PY
