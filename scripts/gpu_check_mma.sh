#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/check_mma.py > gpurun_out/check_mma.log 2>&1; echo "rc=$?" >> gpurun_out/check_mma.log
tail -70 gpurun_out/check_mma.log
