#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 python bench.py --size 4G --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_regrow_4G.json 2> gpurun_out/bench_regrow_4G.err; tail -3 gpurun_out/bench_regrow_4G.err; cut -c1-700 gpurun_out/bench_regrow_4G.json
