#!/bin/bash
# Round-2 first contact: the two window microbenchmarks + the round-1 window path as written.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt 2>&1
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/subpart scripts/micro/subpart.cu && timeout 200 /tmp/subpart > gpurun_out/micro_subpart.txt 2>&1
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/smem_window scripts/micro/smem_window.cu && timeout 200 /tmp/smem_window > gpurun_out/micro_smem_window.txt 2>&1
export JFGPU_K2_WINDOW=1
timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "partition or skew or regrow or grow" > gpurun_out/window_pytest.txt 2>&1
timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/window_bench.txt 2>&1
unset JFGPU_K2_WINDOW
tail -8 gpurun_out/micro_subpart.txt gpurun_out/micro_smem_window.txt; tail -5 gpurun_out/window_pytest.txt
tail -c 1500 gpurun_out/window_bench.txt
