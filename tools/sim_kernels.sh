#!/bin/bash
# per-kernel durations of one SimOTA call (cfg3) under ncu (serialised: PDL overlap is not visible here)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:simota -s 9 -c 3 --csv python tools/sim_time.py 2>/dev/null | grep -E "simota" | awk -F'","' '{print $5, $NF}'
