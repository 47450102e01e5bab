#!/bin/bash
# Round 2: matcher geometries A/B (identical keys vs the INT kernel + timings).  gpurun -- 'bash scripts/gpu_r02_i.sh "0 3"'
set -u
mkdir -p gpurun_out
for G in ${1:-0 1 2 3}; do
  echo "=== SLAMFE_MMA_GEOMETRY=$G"
  SLAMFE_MMA_GEOMETRY=$G timeout 300 python scripts/check_mma.py > gpurun_out/r02i_check_g$G.log 2>&1; echo "rc=$?"
  grep -c "^OK" gpurun_out/r02i_check_g$G.log; grep -E "FAIL|ALL|SOME|mma|Error|error|diffs" gpurun_out/r02i_check_g$G.log | head -12
done
