#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "partitioned" 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
B="python bench.py --bases 2000000000 --size 4G --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv $B > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"count_kernel|insert_chunks" -s 4 -c 3 -o gpurun_out/prof_part_r01 $B > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
