#!/usr/bin/env python
"""profiles/traffic.json from an ncu launch list.

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file gpurun_out/<name>_traffic.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e
  python scripts/tools/make_traffic.py gpurun_out/<name>_traffic.csv profiles/traffic.json [profiles/<name>_launches.csv]

Per kernel: launches, mean duration, mean DRAM bytes read / written per launch (over ALL launches in the list).
bench.py puts `dram_bytes_per_launch` of the dominant kernel into roofline.traffic.  The optional third argument is a
compact per-launch CSV (kernel, duration_us, dram_read, dram_write) that is small enough to commit."""
import collections
import csv
import json
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
lines = [l for l in open(src) if not l.startswith("==")]
per = collections.OrderedDict()       # launch id -> {name, metrics}
for row in csv.DictReader(lines):
    d = per.setdefault(row["ID"], {"name": re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").strip()})
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    m = row["Metric Name"]
    if m == "gpu__time_duration.sum":
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6, "second": v * 1e6}.get(u, v)
    else:
        v = {"byte": v, "Kbyte": v * 1e3, "Mbyte": v * 1e6, "Gbyte": v * 1e9}.get(u, v)
    d[m] = v
agg = collections.OrderedDict()
for d in per.values():
    base = re.sub(r"<.*", "", d["name"])
    a = agg.setdefault(base, {"launches": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "full_name": d["name"]})
    a["launches"] += 1
    a["us"] += d.get("gpu__time_duration.sum", 0.0)
    a["rd"] += d.get("dram__bytes_read.sum", 0.0)
    a["wr"] += d.get("dram__bytes_write.sum", 0.0)
out = {"source": src, "how": "mean over all launches of each kernel in one ncu pass (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum)"}
for k, a in agg.items():
    n = a["launches"]
    out[k] = {"kernel": a["full_name"], "launches": n, "mean_us": a["us"] / n, "dram_read_per_launch": a["rd"] / n, "dram_write_per_launch": a["wr"] / n,
              "dram_bytes_per_launch": (a["rd"] + a["wr"]) / n, "total_ms": a["us"] / 1e3}
json.dump(out, open(dst, "w"), indent=1)
if len(sys.argv) > 3:
    with open(sys.argv[3], "w") as f:
        f.write("kernel,duration_us,dram_read_bytes,dram_write_bytes\n")
        for d in per.values():
            f.write("%s,%.3f,%d,%d\n" % (re.sub(r"<.*", "", d["name"]), d.get("gpu__time_duration.sum", 0), d.get("dram__bytes_read.sum", 0), d.get("dram__bytes_write.sum", 0)))
for k, a in sorted(agg.items(), key=lambda x: -x[1]["us"]):
    print("%-28s n=%5d total %9.3f ms  mean %9.1f us  dram %8.1f MB/launch" % (k, a["launches"], a["us"] / 1e3, a["us"] / a["launches"], (a["rd"] + a["wr"]) / a["launches"] / 1e6))
