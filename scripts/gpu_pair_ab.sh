#!/bin/bash
# A/B of the two-SM (cta_group::2) tcgen05 matcher against the one-SM kernel in a development build
# (SLAMFE_NVCC_DEFINES=-DSLAMFE_MMA_DEV python -m slamfe.build --force): identical keys vs the INT kernel on the
# shape matrix, then kernel timings, for SLAMFE_MMA_PAIR=1 and 0.   gpurun --timeout 400 -- 'bash scripts/gpu_pair_ab.sh'
set -u
mkdir -p gpurun_out
for P in ${1:-1 0}; do
  echo "=== SLAMFE_MMA_PAIR=$P"
  SLAMFE_MMA_PAIR=$P timeout 150 python scripts/check_mma.py > gpurun_out/pair_check_p$P.log 2>&1; echo "check rc=$?"
  grep -c "^OK" gpurun_out/pair_check_p$P.log; grep -E "FAIL|ALL|SOME|mma|Error|error|diffs" gpurun_out/pair_check_p$P.log | head -24
done
