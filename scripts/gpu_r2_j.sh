#!/bin/bash
# Round 2, step j: K1 ring staging actually selected (P = 1024), large-tile scatter: parity, bench, launch list.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --tb=short > gpurun_out/r2j_pytest.txt 2>&1
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2j_bench.txt 2>&1
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2j_launches.csv $B > gpurun_out/ncu_launch.log 2>&1
tail -8 gpurun_out/r2j_pytest.txt; tail -c 600 gpurun_out/r2j_bench.txt
