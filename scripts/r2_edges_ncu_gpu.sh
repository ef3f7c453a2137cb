#!/bin/bash
# Runs ON THE GPU BOX: launch list and ncu --set full of the edge sort's scatter kernel (after the
# plain run exited 0).
mkdir -p gpurun_out
python scripts/edge_sort_run.py > gpurun_out/r2_edge_sort_plain.log 2>&1 && tail -3 gpurun_out/r2_edge_sort_plain.log &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/r2_launches_edge_sort.csv python scripts/edge_sort_run.py > gpurun_out/r2_ncu_es_list.log 2>&1
echo "launch list rc=$?"
python scripts/edge_sort_run.py > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:es_scatter_kernel -s 5 -c 1 -f \
    -o gpurun_out/r2_es_scatter python scripts/edge_sort_run.py > gpurun_out/r2_ncu_es.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r2_es_scatter.ncu-rep
