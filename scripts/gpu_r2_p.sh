#!/bin/bash
# Round 2, final single-GPU pass: whole GPU suite, every BASELINE config, reference arm, DRAM traffic of every launch, ncu captures.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r2p_gpu.txt 2>&1
timeout 1200 python -m pytest tests -q -m gpu --tb=short > gpurun_out/r2p_pytest.txt 2>&1
timeout 600 python bench.py --steps 4 --warmup 3 --dump > gpurun_out/r2p_bench_k21.txt 2>&1
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2p_traffic.csv $B > gpurun_out/ncu_launch.log 2>&1
cap() {  # name regex skip
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c 1 -o /tmp/$1 -f $B >> gpurun_out/ncu_full.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page details > gpurun_out/$1_details.txt 2>/dev/null
}
cap r2p_k1 extract_kernel 12
cap r2p_insert win_insert2 20
timeout 300 python bench.py --config k31 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2p_bench_k31.txt 2>&1
timeout 300 python bench.py --config k63 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2p_bench_k63.txt 2>&1
timeout 400 python bench.py --config bf --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2p_bench_bf.txt 2>&1
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2p_bench_reference.txt 2>&1
du -sm gpurun_out; tail -12 gpurun_out/r2p_pytest.txt; for f in k21 k31 k63 bf reference; do echo == $f; tail -c 400 gpurun_out/r2p_bench_$f.txt; echo; done
