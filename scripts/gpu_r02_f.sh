#!/bin/bash
# Round 2, visit F: GPU suite, createdb workload.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02f_pytest.log
python bench.py --workload createdb --steps 3 > gpurun_out/r02f_createdb.json 2> gpurun_out/r02f_createdb.err; echo "createdb rc=$?"
python bench.py --workload createdb --steps 2 --frames 4540 --no-cpu-baseline > gpurun_out/r02f_createdb_full.json 2> gpurun_out/r02f_createdb_full.err; echo "createdb full rc=$?"
cut -c1-1400 gpurun_out/r02f_createdb.json; cut -c1-1400 gpurun_out/r02f_createdb_full.json
