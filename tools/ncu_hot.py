#!/usr/bin/env python
"""List the hot SASS instructions of an `ncu --page source --csv --print-source sass` export (stdin or file):
per kernel section, every instruction holding more than --min percent of the stall samples, with its top stall reasons."""
import csv, sys, argparse
ap = argparse.ArgumentParser(); ap.add_argument("file"); ap.add_argument("--min", type=float, default=0.7); ap.add_argument("--window", type=str, default="")
a = ap.parse_args()
rows = list(csv.reader(open(a.file)))
sections = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "hdr": None, "data": []}; sections.append(cur); continue
    if cur is None: continue
    if cur["hdr"] is None: cur["hdr"] = r; continue
    cur["data"].append(r)
for sec in sections:
    hdr = sec["hdr"]; data = sec["data"]
    isrc = hdr.index("Source"); isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed")
    st_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_")]
    tot = sum(int(r[isamp] or 0) for r in data)
    print("==", sec["name"][:100], "samples", tot, "instructions", len(data))
    if a.window:
        lo, hi = map(int, a.window.split(":"))
        for k in range(lo, hi):
            r = data[k]; s = int(r[isamp] or 0)
            st = sorted([(int(r[i] or 0), h[6:]) for i, h in st_cols], reverse=True)[:2]
            print(k, r[iex], s, r[isrc][:90], [x for x in st if x[0]])
        continue
    for k, r in enumerate(data):
        s = int(r[isamp] or 0)
        if s > tot * a.min / 100:
            st = sorted([(int(r[i] or 0), h[6:]) for i, h in st_cols], reverse=True)[:2]
            print(k, r[iex], s, "%.1f%%" % (100 * s / tot), r[isrc][:80], st)
