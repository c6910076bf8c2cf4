#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file x.csv ...).
usage: launch_shares.py x.csv   (cold-cache, serialised launches: compare shares, not absolutes)"""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
agg = collections.OrderedDict()
for r in rows[1:]:
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = d["Kernel Name"].split("(")[0].replace("cpd::", "").replace("void ", "")
    v = float(d["Metric Value"].replace(",", ""))
    v = v / 1000.0 if d["Metric Unit"].startswith("n") else (v * 1000.0 if d["Metric Unit"].startswith("m") else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-56s launches %5d  total %10.1f us  share %5.1f%%  avg %9.1f us" % (k[:56], a[0], a[1], 100 * a[1] / tot, a[1] / a[0]))
print("total %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))
