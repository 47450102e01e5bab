#!/bin/bash
# Round 2, visit G: GPU suite; compute-sanitizer memcheck on a small matcher case; N=1 bench (pcie rates).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02g_pytest.log
cat > /tmp/san.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '.')
import slamfe
from slamfe import ops, synth
rng = np.random.default_rng(0)
q = torch.from_numpy(synth.descriptors(rng, 700)).cuda(); t = torch.from_numpy(synth.descriptors(rng, 900)).cuda()
for cols in (False, True):
    for bo in (False, True):
        ops.hamming_top2(q, t, want_cols=cols, best_only=bo)
torch.cuda.synchronize(); print("san done")
PY
timeout 600 compute-sanitizer --tool memcheck python /tmp/san.py > gpurun_out/r02g_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02g_memcheck.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
txt=open('gpurun_out/r02g_bench.json').read()
d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e'].get('pcie'))
PY
