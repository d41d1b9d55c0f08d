#!/usr/bin/env python3
"""Per-source-line summary of an ncu report: python tools/ncu_lines.py <rep> [top]
(uses `ncu --page source --print-source cuda,sass`; needs -lineinfo builds and --import-source on)."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
c_inst, c_samp = ci["Instructions Executed"], ci["# Samples"]
lines = []
tot_i = tot_s = 0
for r in rows[hi + 1:]:
    if len(r) <= c_inst or r[0] == "": continue
    try:
        n = int(r[c_inst]); s = int(r[c_samp])
    except ValueError:
        continue
    lines.append((n, s, r[0], r[1].strip()[:110]))
    tot_i += n; tot_s += s
print(f"total warp-instructions {tot_i:.3e}  samples {tot_s}")
print("by instructions executed:")
for n, s, ln, src in sorted(lines, reverse=True)[:top]:
    print(f"  {100*n/tot_i:5.1f}% inst {100*s/max(tot_s,1):5.1f}% samp  L{ln:>4}  {src}")
print("by stall samples:")
for n, s, ln, src in sorted(lines, key=lambda x: -x[1])[:top // 2]:
    print(f"  {100*n/tot_i:5.1f}% inst {100*s/max(tot_s,1):5.1f}% samp  L{ln:>4}  {src}")
