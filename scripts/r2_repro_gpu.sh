#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_knn_gpu.py -m gpu -x -q > gpurun_out/r2_pytest_knn.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_pytest_knn.log
P="timeout 300 python scripts/prof_run.py --n 1000000 --reps 2"
export PQH_TC_EARLY=1 PQH_TC_SPIN=1 PQH_TC_MMA=2
$P > gpurun_out/r2_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -s 3 -c 1 -f \
    -o gpurun_out/r2_knn_tc_i8_early2 $P > gpurun_out/r2_ncu_early2.log 2>&1
ls -la gpurun_out/r2_knn_tc_i8_early2.ncu-rep; tail -3 gpurun_out/r2_ncu_early2.log
