#!/bin/bash
# Runs ON THE GPU BOX: parity of the rewritten filter pipeline, then a timing matrix over its
# knobs (filter warps, MMA issuers, locality order) on SIFT1M-shaped data, device-resident.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_knn_gpu.py -m gpu -x -q > gpurun_out/r2_pytest_knn.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2_pytest_knn.log
P="timeout 300 python scripts/prof_run.py --n 1000000 --reps 2"
run() { echo "== $*"; env "$@" $P 2>&1 | grep -E "tc dbg|rows" | tail -2 | python scripts/short_stats.py; }
run PQH_TC_ORDER=0
run PQH_TC_ORDER=1
run PQH_TC_ORDER=0 PQH_TC_EPI=16
run PQH_TC_ORDER=1 PQH_TC_EPI=16
run PQH_TC_ORDER=0 PQH_TC_MMA=2
run PQH_TC_ORDER=1 PQH_TC_MMA=2
run PQH_TC_ORDER=1 PQH_TC_EPI=16 PQH_TC_MMA=2
run PQH_TC_ORDER=0 PQH_TC_DEBUG=1
run PQH_TC_ORDER=0 PQH_TC_DEBUG=1 PQH_TC_MMA=2
run PQH_TC_ORDER=0 PQH_TC_DEBUG=2
run PQH_TC_ORDER=0 PQH_TC_DEBUG=2 PQH_TC_EPI=16
run PQH_TC_I8=0 PQH_TC_ORDER=0
run PQH_TC_I8=0 PQH_TC_ORDER=1
run PQH_TC_I8=0 PQH_TC_ORDER=1 PQH_TC_EPI=16
run PQH_TC_I8=0 PQH_TC_ORDER=1 PQH_TC_EPI=16 PQH_TC_MMA=2
