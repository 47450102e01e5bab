#!/bin/bash
# Runs the tcgen05 probe on a B200 (each case under its own timeout: a wrong descriptor can hang an mbarrier wait).
P=scripts/probe_tcgen05
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
for ts in 0 1; do for swap in 0 1; do
  timeout 60 $P check 0 $ts $swap 64 128 0; echo "rc=$?"
done; done
timeout 60 $P check 0 0 0 512 256 0; echo "rc=$?"
timeout 60 $P check 0 1 0 512 256 0; echo "rc=$?"
timeout 60 $P check 1 0 0 64 128 0; echo "rc=$?"
timeout 60 $P check 1 1 0 64 128 0; echo "rc=$?"
timeout 60 $P check 1 0 0 512 128 1; echo "rc=$?"
timeout 60 $P check 1 1 0 512 256 1; echo "rc=$?"
timeout 120 $P rate; echo "rc=$?"
timeout 60 $P ldtm; echo "rc=$?"
timeout 60 $P redux; echo "rc=$?"
} > gpurun_out/probe_tcgen05.log 2>&1
tail -60 gpurun_out/probe_tcgen05.log
