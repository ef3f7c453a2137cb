#!/bin/bash
# Runs ON THE GPU BOX with a diagnostic build (make -C pq_huffman_b200/csrc EXTRA=-DPQH_TC_PROF
# after touching knn_tc.cu): cycle accounting of the filter warps, printed per search with
# PQH_TC_DEBUG=3.  Never the library that is measured or shipped -- rebuild without EXTRA after.
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" python scripts/prof_run.py --n $N --reps 2 2>&1 | python scripts/short_stats.py | tail -${T:-1}; }
for N in 45000 100000 1000000; do
  echo "#### n=$N"
  T=3 run PQH_TC_DEBUG=3
  T=3 run PQH_TC_DEBUG=3 PQH_TC_FLAT_CELLS=1
  T=3 run PQH_TC_DEBUG=3 PQH_TC_CELLS=512
done
