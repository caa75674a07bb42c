#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "partitioned or chunked or lookup or large or route" 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
for cfg in "32 4" "16 4" "16 8" "16 6"; do set -- $cfg
echo "=== REGION_MB=$1 PGRAB=$2"; JFGPU_REGION_MB=$1 JFGPU_PGRAB=$2 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('VALUE', d['value']/1e9, 'ms', d['ms_per_step'], [ (k['kernel'][:12], round(k['seconds'],4)) for k in d['roofline']['kernels']])"
done
