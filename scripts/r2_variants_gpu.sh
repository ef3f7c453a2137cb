#!/bin/bash
# Runs ON THE GPU BOX: whole GPU suite, then a timing matrix over the filter knobs
# (stage release early / late, MMA issuers, locality order) on SIFT1M-shaped data, then ncu.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_all.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2_pytest_all.log
P="timeout 300 python scripts/prof_run.py --n 1000000 --reps 2"
run() { echo "== $*"; env "$@" $P 2>&1 | grep -E "tc dbg|rows" | tail -2 | python scripts/short_stats.py; }
for early in 0 1; do for order in 0 1; do for mma in 1 2; do
  run PQH_TC_EARLY=$early PQH_TC_ORDER=$order PQH_TC_MMA=$mma
done; done; done
run PQH_TC_I8=0 PQH_TC_EARLY=0 PQH_TC_ORDER=1
run PQH_TC_I8=0 PQH_TC_EARLY=1 PQH_TC_ORDER=1
run PQH_TC_I8=0 PQH_TC_EARLY=1 PQH_TC_ORDER=1 PQH_TC_MMA=2
export PQH_TC_ORDER=1 PQH_TC_EARLY=0 PQH_TC_MMA=1
$P > gpurun_out/r2_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -s 1 -c 1 -f \
    -o gpurun_out/r2_knn_tc_i8_late $P > gpurun_out/r2_ncu_late.log 2>&1
export PQH_TC_ORDER=1 PQH_TC_EARLY=1 PQH_TC_MMA=2
$P > gpurun_out/r2_prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -s 1 -c 1 -f \
    -o gpurun_out/r2_knn_tc_i8_early $P > gpurun_out/r2_ncu_early.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
