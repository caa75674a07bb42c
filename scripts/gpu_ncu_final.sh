#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r01_final.csv $B > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"count_kernel" -s 21 -c 1 -o gpurun_out/prof_r01_final_k1 $B > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"insert_chunks" -s 200 -c 1 -o gpurun_out/prof_r01_final_k2 $B >> gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
