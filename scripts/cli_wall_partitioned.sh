#!/bin/bash
# Wall time of the CLI in the reference's DEFAULT mode (5 split coordinates x 3 boxes = 243
# overlapping blocks) on a SIFT-shaped .fvecs file: native driver (data set resident in HBM,
# blocks named by row ids) vs the reference's own binary linked to libpqh_knn.so (re-reads the
# file once per block).  Usage: cli_wall_partitioned.sh [rows] [reference timeout s]
set -e
R=${GRAFT_REPO_ROOT:-/root/repo}
N=${1:-1000000}
REF_TIMEOUT=${2:-600}
python -c "
import sys; sys.path.insert(0, '$R')
from pq_huffman_b200 import synth, vecs_io
vecs_io.write_xvecs('/tmp/part.fvecs', synth.sift_like($N, 128), 'f')"
cd /tmp
for i in 1 2; do sleep 4
  s=$(date +%s%N); $R/pq_huffman_b200/pqh_compute_nn_fast /tmp/part.fvecs /tmp/po_ 16 --temp-file /tmp/pt.vecs | tail -2 | head -1; e=$(date +%s%N)
  echo "pqh_compute_nn_fast (243 blocks, $N rows, blocks merged on the device) run $i: wall $(( (e - s) / 1000000 )) ms"
done
s=$(date +%s%N); $R/pq_huffman_b200/pqh_compute_nn_fast /tmp/part.fvecs /tmp/ph_ 16 --temp-file /tmp/pth.vecs --host-merge | tail -2 | head -1; e=$(date +%s%N)
echo "pqh_compute_nn_fast --host-merge (243 blocks, $N rows): wall $(( (e - s) / 1000000 )) ms"
cmp /tmp/po_nn_indices.ivecsl /tmp/ph_nn_indices.ivecsl && cmp /tmp/po_nn_dist.fvecsl /tmp/ph_nn_dist.fvecsl && echo "device merge == host merge ($N rows)"
s=$(date +%s%N)
if timeout $REF_TIMEOUT $R/oracle/_ref/compute_nn_fast_gpu /tmp/part.fvecs /tmp/pr_ 16 --temp-file /tmp/pt2.vecs > /dev/null; then
  e=$(date +%s%N)
  echo "reference compute_nn_fast + libpqh_knn.so (243 blocks, $N rows): wall $(( (e - s) / 1000000 )) ms"
  cmp /tmp/po_nn_indices.ivecsl /tmp/pr_nn_indices.ivecsl && cmp /tmp/po_nn_dist.fvecsl /tmp/pr_nn_dist.fvecsl && echo "outputs IDENTICAL ($N rows, partitioned)"
else
  e=$(date +%s%N)
  echo "reference compute_nn_fast + libpqh_knn.so: not finished after $(( (e - s) / 1000000 )) ms (timeout $REF_TIMEOUT s)"
fi
