#!/bin/bash
# Runs ON THE GPU BOX: cell count and cell numbering of the locality order on short and long sweeps.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_knn_gpu.py -x -q -m gpu > gpurun_out/r2_pytest_knn.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_knn.log
run() { echo "== $*"; env "$@" python scripts/prof_run.py --n $N --reps 2 2>&1 | python scripts/short_stats.py | tail -${T:-1}; }
for N in 45000 100000; do
  echo "#### n=$N"
  for C in 512 1024 2048 4096; do
    T=2 run PQH_TC_DEBUG=3 PQH_TC_CELLS=$C
    T=2 run PQH_TC_DEBUG=3 PQH_TC_CELLS=$C PQH_TC_FLAT_CELLS=1
  done
  run PQH_TC_CELLS=2048 PQH_TC_OWN=4
  run PQH_TC_CELLS=2048 PQH_TC_OWN=16
  run PQH_TC_CELLS=4096 PQH_TC_OWN=4
  run PQH_TC_CELLS=4096 PQH_TC_OWN=16
  run PQH_TC_CELLS=4096 PQH_TC_NO_TAIL=1
done
N=1000000
echo "#### n=$N"
T=2 run PQH_TC_DEBUG=3
T=2 run PQH_TC_DEBUG=3 PQH_TC_FLAT_CELLS=1
run PQH_TC_OWN=4
run PQH_TC_OWN=16
