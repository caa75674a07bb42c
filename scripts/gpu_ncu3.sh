#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "partitioned or chunked or lookup or large" 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
B="python bench.py --bases 2000000000 --size 4G --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:"count_kernel" -s 4 -c 1 -o gpurun_out/prof_part_count_r01 $B > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"insert_chunks" -s 40 -c 1 -o gpurun_out/prof_part_insert_r01 $B >> gpurun_out/ncu_full.log 2>&1
echo "=== full bench"; timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tee gpurun_out/bench_full4.json
