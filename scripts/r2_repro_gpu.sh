#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_knn_gpu.py tests/test_golden.py -m gpu -x -q > gpurun_out/r2_pytest_knn.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest_knn.log
P="timeout 300 python scripts/prof_run.py --n 1000000 --reps 3"
run() { echo "== $*"; env "$@" $P 2>&1 | grep -E "tc dbg|rows" | tail -1 | python scripts/short_stats.py; }
run PQH_TC_X=0
run PQH_TC_I8=0
echo "== gist1m"; timeout 600 python scripts/prof_run.py --n 1000000 --d 960 --gen gist_like --reps 2 2>&1 | tail -1 | python scripts/short_stats.py
echo "== deep 2M"; timeout 300 python scripts/prof_run.py --n 2000000 --d 96 --k 33 --gen deep_like --reps 2 2>&1 | tail -1 | python scripts/short_stats.py
echo "== sift100k"; timeout 200 python scripts/prof_run.py --n 100000 --reps 3 2>&1 | tail -1 | python scripts/short_stats.py
echo "== 45k"; timeout 200 python scripts/prof_run.py --n 45000 --reps 3 2>&1 | tail -1 | python scripts/short_stats.py
timeout 1500 python -m pytest tests/test_full_size_gpu.py -m gpu -q -k "gist or deep" > gpurun_out/r2_pytest_full.log 2>&1; echo "pytest full rc=$?"; tail -4 gpurun_out/r2_pytest_full.log
