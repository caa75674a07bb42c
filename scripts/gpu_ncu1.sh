#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --bases 1000000000 --size 2G --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv $B > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:count_kernel -s 20 -c 2 -o gpurun_out/prof_count_r01 $B > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
