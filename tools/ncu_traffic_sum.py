#!/usr/bin/env python
"""Sum dram__bytes_read.sum + dram__bytes_write.sum over the launches of ONE call of a multi-kernel operation, from an
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file X` launch list, and record it
in profiles/ncu_traffic.json under KEY (bench.py reads `roofline.traffic` from there).
usage: python tools/ncu_traffic_sum.py X.csv KEY UNITS CALLS [kernel-name regex]
The command profiled made CALLS identical calls of the operation; the launches matching the regex are split into CALLS equal
groups and the LAST group is summed (the first call includes cold-cache / table set-up traffic)."""
import csv
import json
import os
import re
import sys

path, key, units, calls = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
pat = re.compile(sys.argv[5] if len(sys.argv) > 5 else ".")
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
col = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}
launch = {}
for r in rows[rows.index(hdr) + 1:]:
    if not pat.search(r[col["Kernel Name"]]):
        continue
    d = launch.setdefault(r[col["ID"]], {"name": r[col["Kernel Name"]], "bytes": 0.0, "ms": 0.0})
    v = float(r[col["Metric Value"]].replace(",", ""))
    if r[col["Metric Name"]].startswith("dram__bytes"):
        d["bytes"] += v * scale[r[col["Metric Unit"]]]
    elif r[col["Metric Name"]].startswith("gpu__time_duration"):
        d["ms"] += v * scale[r[col["Metric Unit"]]]
ids = sorted(launch, key=int)
per = len(ids) // calls
last = ids[len(ids) - per:]
total = sum(launch[i]["bytes"] for i in last)
ms = sum(launch[i]["ms"] for i in last)
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
d = json.load(open(out)) if os.path.exists(out) else {}
d[key] = {"units": units, "dram_bytes": int(total), "launches": per, "duration_ms_under_ncu": ms, "source": os.path.basename(path)}
json.dump(d, open(out, "w"), indent=1, sort_keys=True)
print(key, "launches per call", per, "dram bytes %.3e" % total, "ms under ncu %.3f" % ms)
