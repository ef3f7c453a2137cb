#!/bin/bash
# Runs ON THE GPU BOX: CLI wall times (both modes) and the GIST1M-shaped evidence -- bench line,
# ncu --set full of the streaming GEMM+filter kernel and of the coalescing re-rank kernel (each
# after the same command exited 0 without ncu).
mkdir -p gpurun_out
bash scripts/cli_wall_1m.sh > gpurun_out/r2_cli_wall_1m.log 2>&1; tail -4 gpurun_out/r2_cli_wall_1m.log
bash scripts/cli_wall_partitioned.sh 1000000 1 > gpurun_out/r2_cli_wall_partitioned.log 2>&1; tail -6 gpurun_out/r2_cli_wall_partitioned.log
G="python bench.py --workload gist1m --steps 1 --warmup 1 --no-cpu --no-ceiling"
python bench.py --workload gist1m --steps 2 --warmup 3 --no-cpu --no-ceiling > gpurun_out/r2_bench_gist1m_n1.json 2> gpurun_out/r2_bench_gist1m_n1.err; echo "gist bench rc=$?"
tail -1 gpurun_out/r2_bench_gist1m_n1.json | cut -c1-600
$G > gpurun_out/r2_bench_gist_short.json 2> gpurun_out/r2_bench_gist_short.err &&
ncu --set full --clock-control none --import-source on -k regex:knn_tcs_kernel -s 4 -c 1 -f \
    -o gpurun_out/r2_knn_tcs_gist1m $G > gpurun_out/r2_ncu_tcs.log 2>&1
echo "ncu tcs rc=$?"
$G > gpurun_out/r2_bench_gist_short2.json 2> gpurun_out/r2_bench_gist_short2.err &&
ncu --set full --clock-control none --import-source on -k regex:knn_rerank_kernel -s 2 -c 1 -f \
    -o gpurun_out/r2_knn_rerank_gist1m $G > gpurun_out/r2_ncu_rerank.log 2>&1
echo "ncu rerank rc=$?"; ls -la gpurun_out/r2_knn_tcs_gist1m.ncu-rep gpurun_out/r2_knn_rerank_gist1m.ncu-rep
