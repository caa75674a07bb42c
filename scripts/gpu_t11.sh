#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "partitioned or chunked or lookup or large" 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
for V in lean generic; do
echo "=== full bench $V"; if [ $V = generic ]; then export JFGPU_K2_GENERIC=1; fi
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tee gpurun_out/bench_full11_$V.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('VALUE', d['value']/1e9, 'ms', d['ms_per_step'], 'feed', d['device_seconds_per_step'], [ (k['kernel'][:20], k['seconds']) for k in d['roofline']['kernels']])"
done
