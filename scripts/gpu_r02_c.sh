#!/bin/bash
# Round 2, visit C: tcgen05 matcher — parity vs the INT kernel on the shape matrix, timings.
set -u
mkdir -p gpurun_out
timeout 300 python scripts/check_mma.py > gpurun_out/r02c_check.log 2>&1; echo "rc=$?"; grep -c "^OK" gpurun_out/r02c_check.log; grep -E "FAIL|ALL|SOME|dense|stereo|Error|error|diffs" gpurun_out/r02c_check.log | head -60
