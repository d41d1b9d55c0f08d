#!/usr/bin/env python3
"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X ...`):
python tools/launch_summary.py launches.csv "<command that was profiled>"  -> per kernel family / per kernel shares."""
import collections, csv, re, sys
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
rd = csv.DictReader(rows)
fam = collections.defaultdict(float); ker = collections.defaultdict(lambda: [0, 0.0]); n = 0; tot = 0.0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum": continue
    v = float(r["Metric Value"].replace(",", ""))
    u = r.get("Metric Unit", "ns")
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    name = r["Kernel Name"]
    m = re.search(r"(eri_\w+|k_\w+)", name)
    base = m.group(1) if m else name.split("(")[0][:40]
    t = re.search(r"<([^>]*)>", name)
    short = base
    if t and base.startswith("eri_"):
        nums = re.findall(r"\d+", t.group(1))
        short = base + "<" + "".join(nums[:4]) + ">"
    fam[base] += v; ker[short][0] += 1; ker[short][1] += v; n += 1; tot += v
print(f"# launch list summary: `ncu --metrics gpu__time_duration.sum --clock-control none` over `{sys.argv[2] if len(sys.argv) > 2 else '?'}`")
print(f"# (cold-cache, serialised: compare shares, not absolute times)  total device time {tot:.1f} ms over {n} launches\n")
print("## by kernel family")
for k, v in sorted(fam.items(), key=lambda kv: -kv[1]): print(f"{k:28s} {v:10.2f} ms {100 * v / tot:5.1f}%")
print("\n## top 40 kernels")
for k, (c, v) in sorted(ker.items(), key=lambda kv: -kv[1][1])[:40]: print(f"{k:34s} n={c:4d} {v:10.3f} ms {100 * v / tot:5.1f}%")
