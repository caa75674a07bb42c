#!/bin/bash
# Round-2 first contact for the experimental window form of K2 (jellyfish_b200/csrc/jf_window.cuh):
#   gpurun --timeout 1500 -- bash scripts/gpu_window.sh
# 1. the two microbenchmarks that model it; 2. parity of the partitioned tests with JFGPU_K2_WINDOW=1
# (under compute-sanitizer for the smallest one); 3. the bench with and without it.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/subpart scripts/micro/subpart.cu && timeout 300 /tmp/subpart > gpurun_out/micro_subpart.txt 2>&1
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/smem_window scripts/micro/smem_window.cu && timeout 300 /tmp/smem_window > gpurun_out/micro_smem_window.txt 2>&1
export JFGPU_K2_WINDOW=1
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "partition or skew or regrow or grow" > gpurun_out/window_pytest.txt 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/window_bench.txt 2>&1
unset JFGPU_K2_WINDOW
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/l2_bench.txt 2>&1
tail -3 gpurun_out/micro_subpart.txt gpurun_out/micro_smem_window.txt gpurun_out/window_pytest.txt
tail -c 600 gpurun_out/window_bench.txt; echo; tail -c 600 gpurun_out/l2_bench.txt
