#!/bin/bash
# Runs ON THE GPU BOX (gpurun): the round-2 evidence set -- whole GPU suite, bench lines, the ncu
# launch list of the bench command and the --set full capture of the headline kernel (each ncu
# pass after the same command exited 0 without ncu; nothing printed under ncu is a bench value).
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_all.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest_all.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_sift1m_n1.json 2> gpurun_out/r2_bench_sift1m_n1.err; echo "bench rc=$?"
tail -1 gpurun_out/r2_bench_sift1m_n1.json | cut -c1-400
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-ceiling"
$B > gpurun_out/r2_bench_short.json 2> gpurun_out/r2_bench_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2_launches_bench_sift1m.csv $B > gpurun_out/r2_ncu_launch.log 2>&1
echo "launch list rc=$?"
$B > gpurun_out/r2_bench_short2.json 2> gpurun_out/r2_bench_short2.err &&
ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -s 5 -c 1 -f \
    -o gpurun_out/r2_knn_tc_i8_sift1m $B > gpurun_out/r2_ncu_i8.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/r2_knn_tc_i8_sift1m.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:knn_rerank_kernel -s 2 -c 1 -f \
    -o gpurun_out/r2_knn_rerank_sift1m $B > gpurun_out/r2_ncu_rerank_sift.log 2>&1
echo "ncu rerank rc=$?"
python bench.py --workload sift100k --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench_sift100k_n1.json 2>/dev/null; tail -1 gpurun_out/r2_bench_sift100k_n1.json | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_n1.json 2>/dev/null; tail -1 gpurun_out/r2_bench_reference_n1.json | cut -c1-300
bash scripts/cli_wall_1m.sh > gpurun_out/r2_cli_wall_1m.log 2>&1; tail -5 gpurun_out/r2_cli_wall_1m.log | cut -c1-600
sleep 5; bash scripts/cli_wall_partitioned.sh 1000000 1 > gpurun_out/r2_cli_wall_partitioned.log 2>&1; tail -8 gpurun_out/r2_cli_wall_partitioned.log | cut -c1-600
