#!/bin/bash
# First GPU contact: tiny parity check against the reference binary built in oracle/_ref.
set -x
cd "$(dirname "$0")/.."
W=${W:-/tmp/jfw}; mkdir -p $W gpurun_out
R=oracle/_ref
nvidia-smi -L
$R/generate_sequence -o $W/seq1m -s 1040104553 1000000 1000000
export SOURCE_DATE_EPOCH=0
for cfg in "21 2M -C" "21 2M" "15 4M -C" "31 2M -C" "40 2M" "63 3M -C" "10 1M -C" "5 1k -C"; do
  set -- $cfg; k=$1; s=$2; c=$3
  $R/jellyfish count -m $k -s $s -t 4 $c -o $W/ref_$k.jf $W/seq1m_0.fa || exit 1
  timeout 120 jellyfish_b200/lib/jellyfish-b200 count -m $k -s $s $c --timing $W/t.txt -o $W/our_$k.jf $W/seq1m_0.fa; rc=$?
  echo "k=$k s=$s $c rc=$rc"; cat $W/t.txt
  python - <<PY
import sys; sys.path.insert(0,'tests')
from jfutil import *
h1,b1=split_db("$W/ref_$k.jf"); h2,b2=split_db("$W/our_$k.jf")
print("RESULT k=$k s=$s $c header_equal", semantic(h1)==semantic(h2), "body_equal", b1==b2, len(b1), len(b2), md5(b1), md5(b2))
if semantic(h1)!=semantic(h2):
    for k_ in SEMANTIC_KEYS:
        if h1.get(k_)!=h2.get(k_): print("  DIFF", k_, str(h1.get(k_))[:200], "|", str(h2.get(k_))[:200])
if b1!=b2:
    r1=dict(records(h1,b1)); r2=dict(records(h2,b2))
    print("  nrec", len(r1), len(r2), "common", len(set(r1)&set(r2)), "count-mismatch", sum(1 for x in r1 if x in r2 and r1[x]!=r2[x]))
    l1=records(h1,b1); l2=records(h2,b2)
    print("  same set/order?", sorted(l1)==sorted(l2), l1[:3], l2[:3])
PY
done 2>&1 | tee gpurun_out/first_contact.log
