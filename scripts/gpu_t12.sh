#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "route_and_shards" 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
