#!/bin/bash
# Runs ON THE GPU BOX: how the number of cells / dimensions of the locality order moves the
# slow-path share and the kernel time (SIFT1M-shaped, device-resident), and the GIST1M test.
mkdir -p gpurun_out
P="timeout 300 python scripts/prof_run.py --n 1000000 --reps 2"
run() { echo "== $*"; env "$@" $P 2>&1 | grep -E "tc dbg|rows" | tail -2 | python scripts/short_stats.py; }
for cells in 1024 2048 4096; do for dims in 32 128; do
  run PQH_TC_CELLS=$cells PQH_TC_ODIMS=$dims PQH_TC_DEBUG=3 PQH_TC_I8=0
  run PQH_TC_CELLS=$cells PQH_TC_ODIMS=$dims
done; done
run PQH_TC_CELLS=4096 PQH_TC_ODIMS=128 PQH_TC_EARLY=1 PQH_TC_MMA=2
run PQH_TC_CELLS=4096 PQH_TC_ODIMS=128 PQH_TC_MMA=2
timeout 900 python -m pytest tests/test_full_size_gpu.py -m gpu -q -k gist > gpurun_out/r2_pytest_gist.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gist.log
