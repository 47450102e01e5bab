#!/bin/bash
# A/B of the persistent tcgen05 matcher against the one-job-per-CTA kernel (development build, see gpu_persist.sh):
# identical keys vs the INT kernel on the shape matrix, then kernel timings, for SLAMFE_MMA_PERSISTENT=1 and 0.
set -u
mkdir -p gpurun_out
for P in 1 0; do
  echo "=== SLAMFE_MMA_PERSISTENT=$P"
  SLAMFE_MMA_PERSISTENT=$P timeout 150 python scripts/check_mma.py > gpurun_out/persist_check_p$P.log 2>&1; echo "check rc=$?"
  grep -c "^OK" gpurun_out/persist_check_p$P.log; grep -E "FAIL|ALL|SOME|mma|Error|error|diffs" gpurun_out/persist_check_p$P.log | head -14
done
