#!/bin/bash
# Round 2, step b: second-generation K1 (extract_kernel) + lane-refill window insert.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
python - > gpurun_out/r2b_sanitizer.txt 2>&1 <<'PY'
import subprocess, sys, os
sys.path.insert(0, "tests")
import gen
os.makedirs("/tmp/sz", exist_ok=True)
open("/tmp/sz/a.fa", "wb").write(gen.fasta(gen._seq(300000, 5)))
r = subprocess.run(["timeout", "300", "compute-sanitizer", "--tool", "memcheck", "jellyfish_b200/lib/jellyfish-b200", "count", "-m", "21", "-s", "1M", "-C", "-o", "/tmp/sz/a.jf", "/tmp/sz/a.fa"], capture_output=True, text=True)
print(r.stdout[-2000:], r.stderr[-2000:])
PY
timeout 900 python -m pytest tests -q -m gpu --tb=line > gpurun_out/r2b_pytest.txt 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench.txt 2>&1
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2b_launches.csv $B > gpurun_out/ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"extract_kernel" -s 21 -c 1 -o gpurun_out/r2b_k1 $B > gpurun_out/ncu_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"win_insert" -s 40 -c 1 -o gpurun_out/r2b_insert $B >> gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/r2b_sanitizer.txt; tail -15 gpurun_out/r2b_pytest.txt; tail -c 600 gpurun_out/r2b_bench.txt
