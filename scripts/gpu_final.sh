#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu --durations=8 2>&1 | tail -16 | tee gpurun_out/pytest_gpu_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "=== default bench"; python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err; cut -c1-400 gpurun_out/bench_default.json
