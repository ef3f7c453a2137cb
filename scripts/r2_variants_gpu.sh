#!/bin/bash
# Runs ON THE GPU BOX: parity of every knob, then the timing matrix on SIFT1M-shaped data.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_knn_gpu.py -m gpu -x -q > gpurun_out/r2_pytest_knn.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2_pytest_knn.log
P="timeout 300 python scripts/prof_run.py --n 1000000 --reps 2"
run() { echo "== $*"; env "$@" $P 2>&1 | grep -E "tc dbg|rows" | tail -2 | python scripts/short_stats.py; }
run PQH_TC_DEBUG=3 PQH_TC_I8=0
run PQH_TC_EARLY=0
run PQH_TC_EARLY=0 PQH_TC_SPIN=1
run PQH_TC_EARLY=0 PQH_TC_SPIN=1 PQH_TC_MMA=2
run PQH_TC_EARLY=1
run PQH_TC_EARLY=1 PQH_TC_SPIN=1
run PQH_TC_EARLY=1 PQH_TC_MMA=2
run PQH_TC_EARLY=1 PQH_TC_SPIN=1 PQH_TC_MMA=2
run PQH_TC_EARLY=1 PQH_TC_SPIN=1 PQH_TC_MMA=2 PQH_TC_CELLS=2048
run PQH_TC_I8=0 PQH_TC_EARLY=1 PQH_TC_SPIN=1 PQH_TC_MMA=2
run PQH_TC_I8=0 PQH_TC_EARLY=0 PQH_TC_SPIN=1
echo "== sift100k"; PQH_TC_EARLY=1 PQH_TC_SPIN=1 PQH_TC_MMA=2 timeout 200 python scripts/prof_run.py --n 100000 --reps 3 2>&1 | tail -1 | python scripts/short_stats.py
echo "== sift100k late"; timeout 200 python scripts/prof_run.py --n 100000 --reps 3 2>&1 | tail -1 | python scripts/short_stats.py
echo "== 45k block"; PQH_TC_EARLY=1 PQH_TC_SPIN=1 PQH_TC_MMA=2 timeout 200 python scripts/prof_run.py --n 45000 --reps 3 2>&1 | tail -1 | python scripts/short_stats.py
echo "== deep 2M"; PQH_TC_EARLY=1 PQH_TC_SPIN=1 PQH_TC_MMA=2 timeout 300 python scripts/prof_run.py --n 2000000 --d 96 --k 33 --gen deep_like --reps 2 2>&1 | tail -1 | python scripts/short_stats.py
echo "== deep 2M late"; timeout 300 python scripts/prof_run.py --n 2000000 --d 96 --k 33 --gen deep_like --reps 2 2>&1 | tail -1 | python scripts/short_stats.py
