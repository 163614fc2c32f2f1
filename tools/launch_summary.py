#!/usr/bin/env python
"""Aggregate an ncu --metrics gpu__time_duration.sum launch list (csv) per kernel."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if r and r[0] == 'ID':
        hdr = r; start = i; break
ik = hdr.index('Kernel Name'); iv = hdr.index('Metric Value')
agg = {}
for r in rows[start + 2:]:
    if len(r) <= iv: continue
    k = r[ik][:70]; v = float(r[iv].replace(',', ''))
    a = agg.setdefault(k, [0, 0]); a[0] += v; a[1] += 1
tot = sum(a[0] for a in agg.values())
print("total %.3f ms over %d launches" % (tot / 1e6, sum(a[1] for a in agg.values())))
for k, (v, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-72s n=%4d total %8.3f ms  avg %8.1f us  %5.1f%%" % (k, n, v / 1e6, v / n / 1e3, 100 * v / tot))
