#!/bin/bash
# Round 2, step l (2 GPUs): record exchange -- emulated on one GPU, then over NCCL; bench at N=1 and N=2; -Q on the device.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/r2l_gpus.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --tb=short -k "record_exchange or partition or skew or route_and_shards or shard_records" > gpurun_out/r2l_pytest_emul.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short -k "quality" > gpurun_out/r2l_pytest_qual.txt 2>&1
if grep -q " failed" gpurun_out/r2l_pytest_emul.txt; then tail -30 gpurun_out/r2l_pytest_emul.txt; tail -30 gpurun_out/r2l_pytest_qual.txt; exit 0; fi
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r2l_bench_n1.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x --tb=short > gpurun_out/r2l_pytest_multi.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2l_bench_n2.txt 2>&1
tail -5 gpurun_out/r2l_pytest_emul.txt; tail -25 gpurun_out/r2l_pytest_qual.txt; tail -12 gpurun_out/r2l_pytest_multi.txt; tail -c 500 gpurun_out/r2l_bench_n1.txt; echo; tail -c 1800 gpurun_out/r2l_bench_n2.txt
