#!/bin/bash
# Round 2, step f: staged K1 tail + TMA double-buffered window insert: parity subset, bench, launch list.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --tb=short -k "partition or skew or grow or large_table or python_api or baseline_config0 or (golden and not corner)" > gpurun_out/r2f_pytest.txt 2>&1
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r2f_bench.txt 2>&1
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2f_launches.csv $B > gpurun_out/ncu_launch.log 2>&1
tail -8 gpurun_out/r2f_pytest.txt; tail -c 900 gpurun_out/r2f_bench.txt
