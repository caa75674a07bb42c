#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_multi.log
echo "=== N=2 bench (small)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 2 --steps 2 --warmup 1 --bases 2000000000 --size 4G > gpurun_out/bench_n2_small.log 2>&1; grep -E "Error|error|^\{" gpurun_out/bench_n2_small.log | tail -5 | cut -c1-800
echo "=== N=2 bench (full)"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_n2_full.log 2>&1; grep -E "Error|error|^\{" gpurun_out/bench_n2_full.log | tail -5 | cut -c1-800
