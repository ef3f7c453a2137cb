#!/bin/bash
# Runs ON A MULTI-GPU BOX (gpurun --gpus N): the in-library NCCL path -- the tests that need two
# devices, then the bench contract line at N GPUs (torchrun, as the driver launches it).
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_knn_gpu.py tests/test_host_cli_gpu.py -m gpu -q \
    -k "multi_gpu or two_gpus" > gpurun_out/r2_pytest_multigpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r2_pytest_multigpu.log
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 \
    > gpurun_out/r2_bench_sift1m_n$N.json 2> gpurun_out/r2_bench_sift1m_n$N.err; echo "bench rc=$?"
tail -1 gpurun_out/r2_bench_sift1m_n$N.json | cut -c1-1500
grep -c "NCCL INFO" gpurun_out/r2_bench_sift1m_n$N.err; grep -m3 "nranks" gpurun_out/r2_bench_sift1m_n$N.err | cut -c1-200
tail -5 gpurun_out/r2_bench_sift1m_n$N.err | cut -c1-300
