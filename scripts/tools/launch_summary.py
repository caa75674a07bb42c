#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, collections, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = re.sub(r'\(.*', '', row['Kernel Name'])[:60]
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v = v / 1e3 if u == 'us' else v / 1e6 if u == 'ns' else v * 1e3 if u in ('s', 'second') else v
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-62s n=%5d total=%9.3f ms  avg=%8.4f ms  %5.1f%%" % (k, n, t, t / n, 100 * t / tot))
print('total %.3f ms' % tot)
