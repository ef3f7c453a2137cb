#!/bin/bash
# Runs ON THE GPU BOX: ncu launch list (per-kernel device time) of the bench command.
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-ceiling"
$B > gpurun_out/r2_bench_short.json 2> gpurun_out/r2_bench_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2_launches_bench_sift1m.csv $B > gpurun_out/r2_ncu_launch.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2_bench_short.json | cut -c1-300
