#!/bin/bash
# Round 2, step h: tile-sorted dump + lean window insert: parity, bench, micro-benchmarks, ncu captures of K1 and the window insert.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/scatter_store scripts/micro/scatter_store.cu && /tmp/scatter_store > gpurun_out/micro_scatter_store.txt 2>&1
timeout 700 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short -k "not large_table" > gpurun_out/r2h_pytest.txt 2>&1
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r2h_bench.txt 2>&1
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
cap() {  # name regex skip
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c 1 -o /tmp/$1 -f $B >> gpurun_out/ncu_full.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page details > gpurun_out/$1_details.txt 2>/dev/null
}
cap r2h_k1 extract_kernel 12
cap r2h_insert win_insert2 20
du -sm gpurun_out; cat gpurun_out/micro_scatter_store.txt; tail -8 gpurun_out/r2h_pytest.txt; tail -c 700 gpurun_out/r2h_bench.txt
