#!/usr/bin/env python3
"""Aggregate an ncu source page by line ranges: python tools/ncu_regions.py <rep> name:lo-hi[,lo-hi] ..."""
import csv, subprocess, sys
rep = sys.argv[1]
regs = []
for a in sys.argv[2:]:
    n, r = a.split(":")
    regs.append((n, [tuple(int(x) for x in p.split("-")) for p in r.split(",")]))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
ci = {h: i for i, h in enumerate(rows[hi])}
c_inst, c_samp = ci["Instructions Executed"], ci["# Samples"]
c_tinst = ci.get("Thread Instructions Executed")
agg = {n: [0, 0, 0] for n, _ in regs}; agg["other"] = [0, 0, 0]
for r in rows[hi + 1:]:
    try:
        ln = int(r[0]); n = int(r[c_inst]); s = int(r[c_samp]); t = int(r[c_tinst]) if c_tinst else 0
    except (ValueError, IndexError):
        continue
    key = "other"
    for name, rr in regs:
        if any(lo <= ln <= hi2 for lo, hi2 in rr): key = name; break
    agg[key][0] += n; agg[key][1] += s; agg[key][2] += t
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
for k, v in agg.items():
    print(f"  {k:14s} {100*v[0]/ti:5.1f}% inst  {100*v[1]/max(ts,1):5.1f}% samples   lanes/inst {v[2]/max(v[0],1):5.1f}")
