#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_multi.log
echo "=== N=2 bench (full)"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus 2 --steps 2 --warmup 2 > gpurun_out/bench_n2_full.log 2>&1; grep -E "rror|^\{" gpurun_out/bench_n2_full.log | tail -5 | cut -c1-1800
