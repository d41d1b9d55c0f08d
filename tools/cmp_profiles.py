#!/usr/bin/env python3
"""Compare per-class times of class_profile outputs: python tools/cmp_profiles.py a.txt b.txt [c.txt ...]"""
import re, sys
L = {'s': 0, 'p': 1, 'd': 2, 'f': 3}
nc = lambda l: (l + 1) * (l + 2) // 2
def load(f):
    d = {}
    for l in open(f):
        m = re.match(r"\s+\((\w\w)\|(\w\w)\)\s+([\d.]+) ms", l)
        if m and ' b' not in l[:20]: d[m.group(1) + m.group(2)] = float(m.group(3))
    return d
tabs = [load(f) for f in sys.argv[1:]]
tot = [0.0] * len(tabs); fam = {}
for k in sorted(tabs[0], key=lambda k: -tabs[0][k]):
    n = 1
    for ch in k: n *= nc(L[ch])
    f = 'small' if n <= 36 else ('medium' if n <= 150 else 'group')
    vals = [t.get(k, float('nan')) for t in tabs]
    for i, v in enumerate(vals):
        tot[i] += v; fam.setdefault(f, [0.0] * len(tabs))[i] += v
    print(f"{k} {f:6s} " + " ".join(f"{v:8.2f}" for v in vals) + "  " + " ".join(f"{vals[0]/v:5.2f}" for v in vals[1:]))
for f, v in fam.items(): print(f"{f:11s} " + " ".join(f"{x:8.1f}" for x in v))
print("total       " + " ".join(f"{x:8.1f}" for x in tot))
