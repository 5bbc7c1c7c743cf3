#!/usr/bin/env python
"""Bucket the warp-stall samples of an `ncu --page source --csv --print-source cuda,sass` dump by kernel phase: the SASS
between two consecutive BAR.SYNC instructions (address order).  usage: python tools/ncu_phases.py src.csv.gz [top_n]"""
import csv
import gzip
import sys

path = sys.argv[1]
rows = list(csv.reader(gzip.open(path, "rt") if path.endswith(".gz") else open(path)))
hdr = next(r for r in rows if r and r[0] == "Line No")
col = {h: i for i, h in enumerate(hdr)}  # first occurrence wins for duplicated names
ia = hdr.index("Address")
isass = ia + 1
stall_cols = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ins = {}
for r in rows:
    if len(r) > isass and r[ia].startswith("0x"):
        a = int(r[ia], 16)
        if a not in ins:
            ins[a] = r
phase, cur = [], {"n": 0, "samples": 0, "exec": 0, "st": {}, "top": []}
for a in sorted(ins):
    r = ins[a]
    s = int(r[col["# Samples"]] or 0)
    cur["n"] += 1
    cur["samples"] += s
    cur["exec"] += int(r[col["Instructions Executed"]] or 0)
    for h, i in stall_cols:
        v = int(r[i] or 0) if r[i] not in ("-", "") else 0
        cur["st"][h] = cur["st"].get(h, 0) + v
    cur["top"].append((s, r[isass].strip(), max(((int(r[i] or 0) if r[i] not in ("-", "") else 0, h) for h, i in stall_cols))[1]))
    if "BAR.SYNC" in r[isass]:
        phase.append(cur)
        cur = {"n": 0, "samples": 0, "exec": 0, "st": {}, "top": []}
phase.append(cur)
tot = sum(p["samples"] for p in phase) or 1
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 4
for k, p in enumerate(phase):
    st = sorted(p["st"].items(), key=lambda x: -x[1])[:5]
    print("phase %d: %5d sass, %5.1f%% of samples, warp-inst %.3g | %s" % (k, p["n"], 100.0 * p["samples"] / tot, p["exec"],
          " ".join("%s=%.0f%%" % (h.replace("stall_", ""), 100.0 * v / max(1, p["samples"])) for h, v in st)))
    for s, txt, why in sorted(p["top"], reverse=True)[:topn]:
        print("      %5.2f%%  %-60s %s" % (100.0 * s / tot, txt[:60], why))
