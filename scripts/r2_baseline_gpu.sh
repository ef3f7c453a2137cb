#!/bin/bash
# Runs ON THE GPU BOX (gpurun): round-2 baseline -- parity suite, INT8 / FP16 timings with the
# MMA-only ceilings (PQH_TC_DEBUG=1: the epilogue releases every tile untouched), and the
# ncu --set full capture of the headline INT8 kernel (after the same command exited 0 plain).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest0.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_pytest0.log
P="python scripts/prof_run.py --n 1000000 --reps 2"
for v in "" "PQH_TC_DEBUG=1" "PQH_TC_I8=0" "PQH_TC_I8=0 PQH_TC_DEBUG=1"; do
  echo "== $v"; env $v $P 2>&1 | tail -3
done
$P > gpurun_out/r2_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -s 1 -c 1 -f \
    -o gpurun_out/r2_knn_tc_i8_sift1m $P > gpurun_out/r2_ncu_i8.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/*.ncu-rep | tail -2
