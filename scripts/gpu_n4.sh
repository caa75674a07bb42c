#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29654 bench.py --gpus 4 --steps 1 --warmup 1 --bases 1000000000 --size 2G --no-cpu-baseline > gpurun_out/bench_n4_small.log 2>&1; grep -E "rror|^\{" gpurun_out/bench_n4_small.log | tail -4 | cut -c1-700
