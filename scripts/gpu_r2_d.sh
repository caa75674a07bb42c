#!/bin/bash
# Round 2, step d (re-entry): state of the tree on a B200 -- full GPU suite, bench, launch list, ncu captures.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r2d_gpu.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu --tb=short > gpurun_out/r2d_pytest.txt 2>&1
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench.txt 2>&1
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2d_launches.csv $B > gpurun_out/ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"extract_kernel" -s 12 -c 1 -o gpurun_out/r2d_k1 -f $B > gpurun_out/ncu_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"win_insert" -s 20 -c 1 -o gpurun_out/r2d_insert -f $B >> gpurun_out/ncu_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"win_scatter" -s 20 -c 1 -o gpurun_out/r2d_scatter -f $B >> gpurun_out/ncu_full.log 2>&1
tail -15 gpurun_out/r2d_pytest.txt; tail -c 1500 gpurun_out/r2d_bench.txt; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out
