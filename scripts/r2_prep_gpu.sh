#!/bin/bash
# Runs ON THE GPU BOX: parity suite, then prep / re-rank times after the vectorised prep kernels.
timeout 900 python -m pytest tests/test_knn_gpu.py -x -q -m gpu > gpurun_out/r2_pytest_knn.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_knn.log
run() { echo "== $*"; env "$@" timeout 300 python scripts/prof_run.py --n $N --d $D --k $K --gen $G --reps 3 2>&1 | python scripts/short_stats.py | tail -2; }
N=1000000 D=128 K=17 G=sift_like
run PQH_TC_I8=1
run PQH_TC_I8=0
N=2000000 D=96 K=33 G=deep_like
run PQH_TC_I8=1
N=200000 D=960 K=17 G=gist_like
run PQH_TC_I8=1
