#!/bin/bash
# Runs ON THE GPU BOX: ncu --set full captures of the two other kernels that carry time --
# knn_tcs_kernel (streaming GEMM+filter, GIST1M-shaped) and knn_rerank_kernel -- each after the
# same command has exited 0 without ncu.  GIST-shaped searches launch each kernel twice (second
# tensor-core pass on the rows whose certificate failed), so "-s 2" picks the main launch of the
# second search.
mkdir -p gpurun_out
G="python bench.py --workload gist1m --steps 1 --warmup 1 --no-cpu"
$G > gpurun_out/bench_gist_short.json 2> gpurun_out/bench_gist_short.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:knn_tcs_kernel -s 2 -c 1 -f \
    -o gpurun_out/r1_knn_tcs_gist1m $G > gpurun_out/ncu_tcs.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_rerank_kernel -s 2 -c 1 -f \
    -o gpurun_out/r1_knn_rerank_gist1m $G > gpurun_out/ncu_rerank_gist.log 2>&1
if [ "$1" = "sift" ]; then
S="python bench.py --steps 1 --warmup 1 --no-cpu"
$S > gpurun_out/bench_sift_short.json 2> gpurun_out/bench_sift_short.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:knn_rerank_kernel -s 1 -c 1 -f \
    -o gpurun_out/r1_knn_rerank_sift1m $S > gpurun_out/ncu_rerank.log 2>&1
fi
ls -la gpurun_out/*.ncu-rep
