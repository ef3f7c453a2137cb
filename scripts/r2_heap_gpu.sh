#!/bin/bash
# Runs ON THE GPU BOX: parity suite of the kNN entry points, then short and long sweeps with the
# heap flavour of the filter on and off.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_knn_gpu.py -x -q -m gpu > gpurun_out/r2_pytest_knn.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_knn.log
run() { echo "== $*"; env "$@" python scripts/prof_run.py --n $N --reps 2 2>&1 | python scripts/short_stats.py | tail -${T:-1}; }
for N in 45000 100000 1000000; do
  echo "#### n=$N"
  T=2 run PQH_TC_DEBUG=3
  T=2 run PQH_TC_DEBUG=3 PQH_TC_HEAP=0
  run PQH_TC_HEAP=1
  run PQH_TC_HEAP=0
  run PQH_TC_CELLS=2048
done
bash scripts/cli_wall_partitioned.sh 1000000 1 2>&1 | tail -8
