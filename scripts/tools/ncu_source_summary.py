#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: stall reasons over the kernel and the hottest SASS instructions."""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr = rows[1]; body = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: 0 for s in stalls}; total = 0
def num(x):
    try: return float(x)
    except ValueError: return 0.0
for r in body:
    if len(r) < len(hdr): continue
    total += num(r[ix["# Samples"]])
    for s in stalls: tot[s] += num(r[ix[s]])
print("samples", int(total), " instructions", len(body), " warp-inst executed", int(sum(num(r[ix["Instructions Executed"]]) for r in body if len(r) >= len(hdr))),
      " thread-inst", int(sum(num(r[ix["Thread Instructions Executed"]]) for r in body if len(r) >= len(hdr))))
for s, v in sorted(tot.items(), key=lambda x: -x[1]):
    if v: print("  %-26s %6.1f%%" % (s, 100 * v / max(1, total)))
print("-- hottest instructions")
order = sorted(range(len(body)), key=lambda i: -num(body[i][ix["# Samples"]]) if len(body[i]) >= len(hdr) else 0)
for i in order[:top]:
    r = body[i]
    best = max(stalls, key=lambda s: num(r[ix[s]]))
    print("%5d %6.2f%% %-70s exec=%-9s thr=%-5s %s" % (i, 100 * num(r[ix["# Samples"]]) / max(1, total), r[ix["Source"]].strip()[:70], r[ix["Instructions Executed"]],
                                                  r[ix["Avg. Threads Executed"]], best))
