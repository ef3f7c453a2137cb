#!/bin/bash
# Runs ON THE GPU BOX (through gpurun): the bench lines and ncu captures that
# scripts/summarise_profiles.py turns into profiles/.  ncu passes only run after the same command
# has exited 0 without ncu; nothing printed under ncu is used as a bench value.
TAG=${1:-r1}
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || exit 1
tail -1 gpurun_out/bench_$TAG.json
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu \
    > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -s 1 -c 1 -f \
    -o gpurun_out/${TAG}_knn_tc_sift1m python bench.py --steps 2 --warmup 1 --no-cpu \
    > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/${TAG}_knn_tc_sift1m.ncu-rep
python bench.py --workload sift100k --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_sift100k.json 2>/dev/null; tail -1 gpurun_out/bench_sift100k.json
python bench.py --workload gist1m --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_gist1m.json 2>/dev/null; tail -1 gpurun_out/bench_gist1m.json
