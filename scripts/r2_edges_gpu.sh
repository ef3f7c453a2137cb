#!/bin/bash
# Runs ON THE GPU BOX: the device edge sort -- its tests, then mst_builder wall time with libc
# qsort against the device sorts on a 1M x 16 graph.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_edge_sort_gpu.py -x -q -m gpu > gpurun_out/r2_pytest_edges.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_pytest_edges.log
timeout 600 python scripts/mst_wall.py 2>&1 | tail -8
