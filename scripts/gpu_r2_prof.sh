#!/bin/bash
# Round 2: window insert as the default K2 -- full GPU suite, bench, launch list, ncu captures.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/r2a_pytest.txt 2>&1
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2a_launches.csv $B > gpurun_out/ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"count_kernel" -s 21 -c 1 -o gpurun_out/r2a_k1 $B > gpurun_out/ncu_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"win_scatter" -s 40 -c 1 -o gpurun_out/r2a_scatter $B >> gpurun_out/ncu_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"win_insert" -s 40 -c 1 -o gpurun_out/r2a_insert $B >> gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/r2a_pytest.txt; tail -c 400 gpurun_out/r2a_bench.txt; tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out
