#!/bin/bash
# ncu capture of the matcher alone (development): gpurun --timeout 900 -- 'bash scripts/gpu_prof_matcher.sh R CS'
set -u
R=${1:-2}; CS=${2:-8}
export TUNE_MODES=11
CMD="python scripts/tune_matcher.py 256 $R $CS"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hamming_top2 -s 1 -c 1 -f -o gpurun_out/prof_matcher_R${R}_CS${CS} $CMD > gpurun_out/prof_ncu.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/prof_plain.log
