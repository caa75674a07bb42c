#!/bin/bash
# Round 2, step n (2 GPUs): restage fixed (one chunk per thread): multi-GPU tests, bench at N=2 with the stage times.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --tb=short -k "record_exchange" > gpurun_out/r2n_pytest_emul.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x --tb=short > gpurun_out/r2n_pytest_multi.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2n_bench_n2.txt 2>&1
tail -3 gpurun_out/r2n_pytest_emul.txt; tail -5 gpurun_out/r2n_pytest_multi.txt; tail -c 900 gpurun_out/r2n_bench_n2.txt
