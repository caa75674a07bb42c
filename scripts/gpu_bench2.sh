#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tee gpurun_out/bench_full2.json
