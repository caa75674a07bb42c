#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for G in 1 2 4; do
echo "=== full bench GRAB=$G"; JFGPU_GRAB=$G timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tee gpurun_out/bench_full8_$G.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('VALUE', d['value']/1e9, 'ms', d['ms_per_step'], 'feed', d['device_seconds_per_step'])"
done
