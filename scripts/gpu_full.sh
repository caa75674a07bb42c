#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "=== default bench"; /usr/bin/time -v python bench.py 2> gpurun_out/bench_default.err | tee gpurun_out/bench_default.json | cut -c1-300; grep -E "Elapsed|Maximum resident" gpurun_out/bench_default.err
echo "=== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 | tee gpurun_out/bench_reference.json | cut -c1-300
