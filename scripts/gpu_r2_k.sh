#!/bin/bash
# Round 2, step k: record exchange emulated on one GPU + everything else; ncu captures of K1, window insert, window scatter.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short -k "not large_table" > gpurun_out/r2k_pytest.txt 2>&1
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
cap() {  # name regex skip
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c 1 -o /tmp/$1 -f $B >> gpurun_out/ncu_full.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page details > gpurun_out/$1_details.txt 2>/dev/null
}
cap r2k_k1 extract_kernel 12
cap r2k_insert win_insert2 20
cap r2k_scatter win_scatter 20
du -sm gpurun_out; tail -15 gpurun_out/r2k_pytest.txt
