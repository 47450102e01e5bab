#!/bin/bash
# One GPU visit: parity tests, smoke, default bench (+ reference arm), ncu launch list of the bench command.
# Run as: gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh TAG'
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "bench ref rc=$?"
python - <<PY
import json
txt=open('gpurun_out/${TAG}_bench.json').read()
d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'equal',d['e2e']['tables_equal_resident_run'])
print('roofline',d['roofline']['frac'],d['roofline']['frac_of_mma_issue_floor'],d['roofline']['launch_ms'])
print('parity',d['parity']['match_tables_bit_exact'],d['parity']['mutual_matches_bit_exact'],d['parity']['xyz_max_rel_err'],d['parity']['ransac_vs_ground_truth']['frac_pairs_pose_within_1cm_1mrad'])
for k,v in d['extra'].items(): print(k,v['value'],v['e2e']['value'],v['parity'])
PY
