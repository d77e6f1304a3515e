#!/usr/bin/env python
"""Hot-spot summary of `ncu -i X.ncu-rep --page source --csv --kernel-name regex:K` output:
samples between barriers and the top instructions of the first kernel instance."""
import csv
import sys

r = list(csv.reader(open(sys.argv[1])))
hdr = None
rows = []
ninst = 0
for x in r:
    if x and x[0] == "Address":
        if hdr is not None:
            break
        hdr = x
        continue
    if hdr is not None and len(x) == len(hdr):
        rows.append(x)
si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
tot = sum(int(x[si]) for x in rows)
print("total samples", tot, "instrs", len(rows))
acc, start = 0, 0
for i, x in enumerate(rows):
    acc += int(x[si])
    if "BAR" in x[src] or i == len(rows) - 1:
        if acc > tot * 0.005:
            print("insns %5d-%5d  samples %6d  %5.1f%%  ends with %-28s exec %s" % (start, i, acc, 100.0 * acc / tot, x[src].strip()[:28], x[ie]))
        acc, start = 0, i + 1
top = sorted(range(len(rows)), key=lambda i: -int(rows[i][si]))[: int(sys.argv[2]) if len(sys.argv) > 2 else 20]
for i in sorted(top):
    print(i, rows[i][si], rows[i][ie], rows[i][src].strip()[:100])
