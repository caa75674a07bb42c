#!/bin/bash
# Round 2, step c: per-CTA chunk arenas + all-32-bit tail of K1.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --tb=line > gpurun_out/r2c_pytest.txt 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench.txt 2>&1
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2c_launches.csv $B > gpurun_out/ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"extract_kernel" -s 36 -c 1 -o gpurun_out/r2c_k1 $B > gpurun_out/ncu_full.log 2>&1
tail -8 gpurun_out/r2c_pytest.txt; tail -c 600 gpurun_out/r2c_bench.txt
