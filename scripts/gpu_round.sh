#!/bin/bash
# One GPU visit: parity tests, smoke, bench, ncu launch list, ncu full capture of the matcher.
# Run as: gpurun --timeout 1700 -- 'bash scripts/gpu_round.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/clocks.csv &
SMI=$!
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
kill $SMI
BENCH_SMALL="python bench.py --frames 512 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$BENCH_SMALL > gpurun_out/plain_small.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"hamming|stereo_|triangulate|ransac|track_gather|scatter_inliers|peak_|unpack_keys|merge_top2|cross_check|ratio_test|nccl" -c 400 --csv --log-file gpurun_out/launches.csv $BENCH_SMALL > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
$BENCH_SMALL > gpurun_out/plain_small2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hamming_top2 -s 4 -c 2 -f -o gpurun_out/prof_hamming $BENCH_SMALL > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
# DRAM traffic of the dominant launch at the FULL bench size (one launch, dram metrics only)
BENCH_FULL="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$BENCH_FULL > gpurun_out/plain_full.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:hamming_top2 -s 6 -c 2 --csv --log-file gpurun_out/traffic_full.csv $BENCH_FULL > gpurun_out/ncu_traffic.log 2>&1
echo "ncu traffic rc=$?"
# full captures of the other kernels of a step (scorer, generator, gather, stereo epilogue, triangulation)
$BENCH_SMALL > gpurun_out/plain_small3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"ransac_score|ransac_hypotheses|track_gather|stereo_links|triangulate_links" -s 15 -c 5 -f -o gpurun_out/prof_other $BENCH_SMALL > gpurun_out/ncu_other.log 2>&1
echo "ncu other rc=$?"
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json; cat gpurun_out/bench_ref.json
