#!/bin/bash
# Round 2, step e: launch list + ncu captures exported to CSV on the box (the .ncu-rep files exceed the 64 MiB return limit).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2e_launches.csv $B > gpurun_out/ncu_launch.log 2>&1
cap() {  # name regex skip
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c 1 -o /tmp/$1 -f $B >> gpurun_out/ncu_full.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page details > gpurun_out/$1_details.txt 2>/dev/null
}
cap r2e_k1 extract_kernel 12
cap r2e_insert win_insert 20
cap r2e_scatter win_scatter 20
cap r2e_hist win_hist 20
timeout 300 python scripts/edge_diag.py > gpurun_out/r2e_edge_diag.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "bloom_prefilter" --tb=short > gpurun_out/r2e_pytest_bf.txt 2>&1
du -sm gpurun_out; ls -la gpurun_out; tail -3 gpurun_out/r2e_pytest_bf.txt; cat gpurun_out/r2e_edge_diag.txt
