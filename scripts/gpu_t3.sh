#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "=== small bench"; timeout 300 python bench.py --bases 1000000000 --size 2G --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tee gpurun_out/bench_small3.json
echo "=== full bench"; timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tee gpurun_out/bench_full3.json
