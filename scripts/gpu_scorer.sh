#!/bin/bash
# Scorer iteration: the RANSAC / loop / front-end parity tests, then the loop-closure and sequence-RANSAC workloads.
# gpurun --timeout 900 -- 'bash scripts/gpu_scorer.sh TAG'
set -u
TAG=${1:-scorer}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ransac.py tests/test_gpu_loop.py tests/test_gpu_frontend.py -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
python bench.py --workload loop --steps 5 --warmup 3 > gpurun_out/${TAG}_loop.json 2> gpurun_out/${TAG}_loop.err
python bench.py --workload ransac --steps 5 --warmup 3 > gpurun_out/${TAG}_ransac.json 2> gpurun_out/${TAG}_ransac.err
python - <<PY
import json
for f in ["${TAG}_loop", "${TAG}_ransac"]:
    txt = open(f"gpurun_out/{f}.json").read()
    d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
    print(f, d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d.get("parity"))
PY
