#!/bin/bash
# Persistent tcgen05 matcher check: identical keys vs the INT kernel on the shape matrix + kernel timings, then the
# matcher / front-end GPU tests and the default bench — each step only if the one before passed.
# A/B against the one-job-per-CTA kernel needs a development build (SLAMFE_NVCC_DEFINES=-DSLAMFE_MMA_DEV python -m
# slamfe.build --force) and SLAMFE_MMA_PERSISTENT=0.   gpurun --timeout 900 -- 'bash scripts/gpu_persist.sh'
set -u
mkdir -p gpurun_out
P=${SLAMFE_MMA_PERSISTENT:-1}
echo "=== SLAMFE_MMA_PERSISTENT=$P"
SLAMFE_MMA_PERSISTENT=$P timeout 200 python scripts/check_mma.py > gpurun_out/persist_check_p$P.log 2>&1; rc=$?; echo "check rc=$rc"
grep -c "^OK" gpurun_out/persist_check_p$P.log; grep -E "FAIL|ALL|SOME|mma|Error|error|diffs" gpurun_out/persist_check_p$P.log | head -14
grep -q "ALL OK" gpurun_out/persist_check_p$P.log || exit 1
[ $rc = 0 ] || exit 1
SLAMFE_MMA_PERSISTENT=$P timeout 400 python -m pytest tests/test_gpu_matching.py tests/test_gpu_frontend.py -m gpu -x -q > gpurun_out/persist_pytest_p$P.log 2>&1; rc=$?
echo "pytest rc=$rc"; tail -3 gpurun_out/persist_pytest_p$P.log
[ $rc = 0 ] || exit 1
SLAMFE_MMA_PERSISTENT=$P timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/persist_bench_p$P.json 2> gpurun_out/persist_bench_p$P.err; echo "bench rc=$?"
python - <<PY
import json
try:
    txt=open('gpurun_out/persist_bench_p$P.json').read()
    d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'equal',d['e2e']['tables_equal_resident_run'])
    print('roofline',d['roofline']['frac'],d['roofline']['frac_of_mma_issue_floor'],d['roofline']['launch_ms'])
    print('parity',d['parity']['match_tables_bit_exact'],d['parity']['mutual_matches_bit_exact'])
    for k,v in d['extra'].items(): print(k,v['value'],v['e2e']['value'],v['parity'])
except Exception as e:
    print('no bench line', e)
PY
tail -3 gpurun_out/persist_bench_p$P.err
