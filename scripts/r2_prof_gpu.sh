#!/bin/bash
# Runs ON THE GPU BOX with a -DPQH_TC_PROF build: cycle accounting of the filter warps.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_knn_gpu.py -x -q -m gpu > gpurun_out/r2_pytest_knn.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_knn.log
run() { echo "== $*"; env "$@" python scripts/prof_run.py --n $N --reps 2 2>&1 | python scripts/short_stats.py | tail -${T:-1}; }
for N in 45000 100000 1000000; do
  echo "#### n=$N"
  T=3 run PQH_TC_DEBUG=3 PQH_TC_CELLS=4096
  T=3 run PQH_TC_DEBUG=3 PQH_TC_CELLS=4096 PQH_TC_SOFT=250
  T=3 run PQH_TC_DEBUG=3 PQH_TC_CELLS=4096 PQH_TC_SOFT=250 PQH_TC_LEXACT=20
  run PQH_TC_CELLS=4096
  run PQH_TC_CELLS=4096 PQH_TC_SOFT=200
  run PQH_TC_CELLS=4096 PQH_TC_SOFT=250
  run PQH_TC_CELLS=4096 PQH_TC_SOFT=400
  run PQH_TC_CELLS=4096 PQH_TC_SOFT=250 PQH_TC_LEXACT=20
  run PQH_TC_CELLS=2048 PQH_TC_SOFT=250 PQH_TC_LEXACT=20
done
