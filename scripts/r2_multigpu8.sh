#!/bin/bash
# Runs ON AN 8-GPU BOX (gpurun --gpus 8): multi-device tests, the Deep10M-shaped full-size parity
# test through the library's own 8-GPU path, and the bench contract line at 8 and 4 GPUs (torchrun,
# as the driver launches it).
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_knn_gpu.py tests/test_host_cli_gpu.py -m gpu -q \
    -k "multi_gpu or two_gpus" > gpurun_out/r2_pytest_multigpu8.log 2>&1; echo "pytest multi rc=$?"
tail -3 gpurun_out/r2_pytest_multigpu8.log
timeout 900 python -m pytest tests/test_full_size_gpu.py -m gpu -q -k deep > gpurun_out/r2_pytest_deep10m.log 2>&1; echo "pytest deep10m rc=$?"
tail -3 gpurun_out/r2_pytest_deep10m.log
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 --no-cpu \
      > gpurun_out/r2_bench_sift1m_n$N.json 2> gpurun_out/r2_bench_sift1m_n$N.err; echo "bench N=$N rc=$?"
  tail -1 gpurun_out/r2_bench_sift1m_n$N.json | cut -c1-1200
  tail -3 gpurun_out/r2_bench_sift1m_n$N.err | cut -c1-300
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
    --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --workload deep10m --steps 2 --warmup 3 --no-cpu \
    > gpurun_out/r2_bench_deep10m_n8.json 2> gpurun_out/r2_bench_deep10m_n8.err; echo "bench deep10m rc=$?"
tail -1 gpurun_out/r2_bench_deep10m_n8.json | cut -c1-1200
