#!/bin/bash
# Round 2, last call: experiment -- jfgpu_feed with scripts/experiments/feed_staging_ring.patch applied (not adopted).
# Whole GPU suite first (parity is the gate), then the headline configuration with its e2e leg.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 235 python -m pytest tests -q -m gpu -x --tb=short > gpurun_out/r2s_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2s_pytest.txt
timeout 130 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2s_bench_k21.txt 2>&1
echo "bench rc=$?" >> gpurun_out/r2s_bench_k21.txt
tail -4 gpurun_out/r2s_pytest.txt; tail -c 1500 gpurun_out/r2s_bench_k21.txt
