#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_mma.py mma > gpurun_out/prof_mma_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hamming_mma -s 3 -c 3 -f -o gpurun_out/r02_mma_v1 python scripts/prof_mma.py mma > gpurun_out/prof_mma_ncu.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/prof_mma_ncu.log
