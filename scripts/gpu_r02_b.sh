#!/bin/bash
# Round 2, visit B: GPU suite (both matcher kernels), default bench with sub-records, reference arm, launch list, DRAM traffic.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02b_pytest.log
python __graft_entry__.py smoke > gpurun_out/r02b_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02b_bench_ref.json 2> gpurun_out/r02b_bench_ref.err; echo "bench ref rc=$?"
BENCH_SMALL="python bench.py --frames 512 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches.csv $BENCH_SMALL > gpurun_out/r02b_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
BENCH_FULL="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:hamming_mma -s 8 -c 2 --csv --log-file gpurun_out/r02b_traffic_full.csv $BENCH_FULL > gpurun_out/r02b_ncu_traffic.log 2>&1
echo "ncu traffic rc=$?"
cat gpurun_out/r02b_bench.json; cat gpurun_out/r02b_bench_ref.json
