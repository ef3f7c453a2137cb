#!/bin/bash
# Runs ON THE GPU BOX: A/B of two builds of the library on the same box, interleaved
# (pq_huffman_b200/libpqh_knn.so against pq_huffman_b200/libpqh_knn_$1.so).
V=$1
L=pq_huffman_b200/libpqh_knn.so
cp $L /tmp/base.so
run() { python scripts/prof_run.py --n ${N:-1000000} --reps 3 2>&1 | python scripts/short_stats.py | tail -2; }
for i in 1 2; do
  echo "== base"; cp /tmp/base.so $L; run
  echo "== $V"; cp pq_huffman_b200/libpqh_knn_$V.so $L; run
done
echo "== base 100k"; cp /tmp/base.so $L; N=100000 run
echo "== $V 100k"; cp pq_huffman_b200/libpqh_knn_$V.so $L; N=100000 run
cp /tmp/base.so $L
