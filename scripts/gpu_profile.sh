#!/bin/bash
# Profiles of the shipped build: ncu launch list of the bench command, ncu --set full of the matcher, DRAM traffic of
# the dominant launch at full bench size.  gpurun --timeout 1500 -- 'bash scripts/gpu_profile.sh TAG'
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
BENCH_SMALL="python bench.py --frames 512 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
$BENCH_SMALL > gpurun_out/${TAG}_plain_small.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"hamming|stereo_|triangulate|ransac|track_|scatter_inliers|pnp_refit|pairs_gather|pack_db|link_offsets|gate_|unpack_keys|merge_top2|cross_check|ratio_test|nccl" -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH_SMALL > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python scripts/prof_mma.py mma > gpurun_out/${TAG}_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hamming_mma -s 3 -c 3 -f -o gpurun_out/${TAG}_mma python scripts/prof_mma.py mma > gpurun_out/${TAG}_prof_ncu.log 2>&1
echo "ncu full rc=$?"
BENCH_FULL="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
$BENCH_FULL > gpurun_out/${TAG}_plain_full.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:hamming_mma -s 8 -c 2 --csv --log-file gpurun_out/${TAG}_traffic_full.csv $BENCH_FULL > gpurun_out/${TAG}_ncu_traffic.log 2>&1
echo "ncu traffic rc=$?"
tail -4 gpurun_out/${TAG}_traffic_full.csv | cut -c1-300
