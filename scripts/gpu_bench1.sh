#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L; nproc; free -g | head -2
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "=== small bench"; python bench.py --bases 1000000000 --size 2G --steps 2 --warmup 1 --cpu-sample-bases 100000000 2>&1 | tee gpurun_out/bench_small.json
echo "=== full bench"; python bench.py --steps 3 --warmup 2 2>&1 | tee gpurun_out/bench_full.json
