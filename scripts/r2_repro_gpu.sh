#!/bin/bash
mkdir -p gpurun_out
P="timeout 300 python scripts/prof_run.py --n 1000000 --reps 3"
run() { echo "== $*"; env "$@" $P 2>&1 | grep -E "tc dbg|rows" | tail -1 | python scripts/short_stats.py; }
run PQH_TC_STAGGER=0 PQH_TC_OWN=8
run PQH_TC_STAGGER=2 PQH_TC_OWN=8
run PQH_TC_STAGGER=8 PQH_TC_OWN=8
run PQH_TC_STAGGER=0 PQH_TC_OWN=4
run PQH_TC_STAGGER=0 PQH_TC_OWN=16
run PQH_TC_STAGGER=0 PQH_TC_OWN=32
run PQH_TC_STAGGER=4 PQH_TC_OWN=16
run PQH_TC_STAGGER=0 PQH_TC_OWN=8 PQH_TC_DEBUG=3 PQH_TC_I8=0
run PQH_TC_STAGGER=0 PQH_TC_OWN=32 PQH_TC_DEBUG=3 PQH_TC_I8=0
echo "== gist1m"; timeout 600 python scripts/prof_run.py --n 1000000 --d 960 --gen gist_like --reps 2 2>&1 | tail -2 | python scripts/short_stats.py
timeout 1500 python -m pytest tests/test_full_size_gpu.py -m gpu -q > gpurun_out/r2_pytest_full.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest_full.log
