#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== default bench"; time (python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err); tail -3 gpurun_out/bench_default.err; cut -c1-2500 gpurun_out/bench_default.json
