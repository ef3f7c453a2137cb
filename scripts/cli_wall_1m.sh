#!/bin/bash
# Wall time of the whole CLI on a SIFT1M-shaped .fvecs file: native host driver vs the
# reference's own binary linked to libpqh_knn.so (both write identical bytes).
set -e
R=${GRAFT_REPO_ROOT:-/root/repo}
python -c "
import sys; sys.path.insert(0, '$R')
from pq_huffman_b200 import synth, vecs_io
vecs_io.write_xvecs('/tmp/sift1m.fvecs', synth.sift_like(1000000, 128), 'f')"
cd /tmp
for i in 1 2; do sleep 4
  SECONDS=0; s=$(date +%s%N); $R/pq_huffman_b200/pqh_compute_nn_fast /tmp/sift1m.fvecs /tmp/o_ 16 --num-dims 0 --temp-file /tmp/t.vecs | tail -2 | head -1; e=$(date +%s%N)
  echo "pqh_compute_nn_fast run $i: wall $(( (e - s) / 1000000 )) ms"
done
s=$(date +%s%N); $R/oracle/_ref/compute_nn_fast_gpu /tmp/sift1m.fvecs /tmp/r_ 16 --num-dims 0 --temp-file /tmp/t2.vecs > /dev/null; e=$(date +%s%N)
echo "reference compute_nn_fast + libpqh_knn.so: wall $(( (e - s) / 1000000 )) ms"
cmp /tmp/o_nn_indices.ivecsl /tmp/r_nn_indices.ivecsl && cmp /tmp/o_nn_dist.fvecsl /tmp/r_nn_dist.fvecsl && echo "outputs IDENTICAL at 1M"
