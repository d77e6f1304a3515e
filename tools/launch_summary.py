#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

rows = [l for l in open(sys.argv[1]) if l.startswith('"ID"') or (l.startswith('"') and l[1].isdigit())]
r = list(csv.reader(rows))
hdr, data = r[0], r[1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(list)
for d in data:
    v = float(d[vi].replace(",", ""))
    v = v / 1000 if d[ui] == "ns" else (v * 1000 if d[ui] == "ms" else v)
    agg[d[ki][:70]].append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-72s n=%3d avg %8.1f us  min %8.1f  max %8.1f  share %5.1f%%" % (k, len(v), sum(v) / len(v), min(v), max(v), 100 * sum(v) / tot))
