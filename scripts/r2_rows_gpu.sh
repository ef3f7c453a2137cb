#!/bin/bash
# Runs ON THE GPU BOX: GEMM+filter time against the number of query rows at a fixed 1M-row base.
timeout 900 python -m pytest tests/test_knn_gpu.py -x -q -m gpu > gpurun_out/r2_pytest_knn.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_knn.log
run() { echo "== rows=$R $*"; env "$@" python scripts/prof_run.py --n 1000000 --rows $R --reps 2 2>&1 | python scripts/short_stats.py | tail -${T:-1}; }
for R in 37888 125000 250000 1000000; do
  T=2 run PQH_TC_DEBUG=3
done
R=125000 run PQH_TC_OWN=16
R=125000 run PQH_TC_OWN=4
