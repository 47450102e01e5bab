#!/bin/bash
# Round 2, visit D: GPU suite (both matcher kernels), default bench with sub-records, launch list, ncu full of the matcher.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02d_pytest.log
python __graft_entry__.py smoke > gpurun_out/r02d_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$?"
BENCH_SMALL="python bench.py --frames 512 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"hamming|stereo_|triangulate|ransac|track_gather|scatter_inliers|pnp_refit|pairs_gather|peak_|unpack_keys|merge_top2|cross_check|ratio_test|nccl" -c 200 --csv --log-file gpurun_out/r02d_launches.csv $BENCH_SMALL > gpurun_out/r02d_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python scripts/prof_mma.py mma > gpurun_out/r02d_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hamming_mma -s 3 -c 3 -f -o gpurun_out/r02_mma_v3 python scripts/prof_mma.py mma > gpurun_out/r02d_prof_ncu.log 2>&1
echo "ncu full rc=$?"
cat gpurun_out/r02d_bench.json | cut -c1-1500
