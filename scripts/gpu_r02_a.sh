#!/bin/bash
# Round 2, visit A: whole GPU suite with the tcgen05 matcher forced on, bench both ways, launch list.
set -u
mkdir -p gpurun_out
SLAMFE_MATCH_MMA=1 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest_mma.log 2>&1; echo "pytest(mma) rc=$?"
tail -3 gpurun_out/r02a_pytest_mma.log
SLAMFE_MATCH_MMA=1 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench_mma.json 2> gpurun_out/r02a_bench_mma.err; echo "bench(mma) rc=$?"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_int.json 2> gpurun_out/r02a_bench_int.err; echo "bench(int) rc=$?"
BENCH_SMALL="python bench.py --frames 512 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
SLAMFE_MATCH_MMA=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches_mma.csv $BENCH_SMALL > gpurun_out/r02a_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
cat gpurun_out/r02a_bench_mma.json; cat gpurun_out/r02a_bench_int.json
