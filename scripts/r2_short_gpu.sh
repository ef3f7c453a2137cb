#!/bin/bash
# Runs ON THE GPU BOX: where the time goes on short sweeps (100k rows, one 45k-row block) and in
# the partitioned CLI.
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" python scripts/prof_run.py --n $N --reps 2 2>&1 | python scripts/short_stats.py | tail -${T:-1}; }
for N in 100000 45000; do
  echo "#### n=$N"
  T=2 run PQH_TC_DEBUG=3
  T=2 run PQH_TC_DEBUG=3 PQH_TC_ORDER=0
  run PQH_TC_DEBUG=2
  run PQH_TC_DEBUG=1
  run PQH_TC_OWN=2
  run PQH_TC_OWN=16
  run PQH_TC_OWN=32
  run PQH_TC_CELLS=1024
  run PQH_TC_CELLS=2048
  run PQH_TC_EARLY=0
  run PQH_TC_MMA=1
  run PQH_TC_I8=0
done
bash scripts/cli_wall_partitioned.sh 1000000 1 2>&1 | tail -8
