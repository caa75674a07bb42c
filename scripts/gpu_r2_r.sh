#!/bin/bash
# Round 2, step r (2 GPUs): own chunks skip the exchange -- multi-GPU tests and the bench at N=2.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -q -m gpu -x --tb=short > gpurun_out/r2r_pytest_multi.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r2r_bench_n2.txt 2>&1
tail -5 gpurun_out/r2r_pytest_multi.txt; tail -c 600 gpurun_out/r2r_bench_n2.txt
