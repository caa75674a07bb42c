#!/bin/bash
# tests on the GPU box + keep the log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -x -q -m gpu 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
