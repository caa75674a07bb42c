#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
W=/tmp/cfg1; mkdir -p $W
echo "=== BASELINE configs[0]: 100 Mbp, k=21, -s 100M -C"
oracle/_ref/generate_sequence -o $W/seq100m -s 3141592653 100000000
md5sum $W/seq100m.fa
export SOURCE_DATE_EPOCH=0
( time oracle/_ref/jellyfish count -m 21 -s 100M -t 32 -C --timing $W/ref.t -o $W/ref.jf $W/seq100m.fa ) 2>&1 | grep real; cat $W/ref.t
( time jellyfish_b200/lib/jellyfish-b200 count -m 21 -s 100M -C --timing $W/our.t -o $W/our.jf $W/seq100m.fa ) 2>&1 | grep real; cat $W/our.t
python - <<PY | tee gpurun_out/cfg1_parity.txt
import sys; sys.path.insert(0,'tests')
from jfutil import *
h1,b1=split_db("$W/ref.jf"); h2,b2=split_db("$W/our.jf")
print("cfg1 header_equal", semantic(h1)==semantic(h2), "body_equal", b1==b2, "records", len(b1)//10, "md5 ref", md5(b1), "ours", md5(b2), "survey-pinned 63058a336e1d9431eb6618d4a4f4deed")
PY
for cfg in "4 4" "2 4" "8 4" "4 2" "4 1"; do set -- $cfg
echo "=== PGRAB=$1 LOOK1=$2"; JFGPU_PGRAB=$1 JFGPU_LOOK1=$2 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('VALUE', d['value']/1e9, 'ms', d['ms_per_step'], [ (k['kernel'][:12], round(k['seconds'],4)) for k in d['roofline']['kernels']])"
done
