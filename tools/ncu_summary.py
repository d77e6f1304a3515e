#!/usr/bin/env python
"""Per-kernel summary (JSON) of `ncu -i X.ncu-rep --page raw --csv` output: one entry per kernel name, first instance
after the skipped warm-up launches.  usage: ncu_summary.py raw.csv [raw2.csv ...] > summary.json"""
import csv
import json
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__waves_per_multiprocessor"]
out, seen = [], {}
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        name = r[ki].split("(")[0].replace("plyolo::", "").replace("void ", "")
        seen[name] = seen.get(name, 0) + 1
        if seen[name] != 2 and not (seen[name] == 1 and "--first" in sys.argv):
            continue  # second instance: warm
        e = {"kernel": name}
        for h, u, v in zip(hdr, units, r):
            if h in WANT:
                e[h] = ("%s %s" % (v, u)).strip()
        out = [o for o in out if o["kernel"] != name] + [e]
print(json.dumps(out, indent=1))
