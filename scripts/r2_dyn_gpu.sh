#!/bin/bash
# Runs ON THE GPU BOX: work items on demand (default) against round robin.
timeout 900 python -m pytest tests/test_knn_gpu.py -x -q -m gpu > gpurun_out/r2_pytest_knn.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_knn.log
run() { echo "== n=$N rows=$R $*"; env "$@" timeout 300 python scripts/prof_run.py --n $N --rows $R --reps 3 2>&1 | python scripts/short_stats.py | tail -${T:-2}; }
N=1000000; R=1000000
run PQH_TC_DYN=1
run PQH_TC_DYN=0
run PQH_TC_DYN=1 PQH_TC_I8=0
run PQH_TC_DYN=0 PQH_TC_I8=0
T=3 run PQH_TC_DYN=1 PQH_TC_DEBUG=3
R=125000
run PQH_TC_DYN=1
run PQH_TC_DYN=0
N=100000; R=100000
run PQH_TC_DYN=1
run PQH_TC_DYN=0
N=45000; R=45000
run PQH_TC_DYN=1
run PQH_TC_DYN=0
