#!/usr/bin/env python3
"""Per-class and per-(class, contraction bucket pair) profile of one Fock build (serialised launches).
usage: python tools/class_profile.py <workload> [out.txt]"""
import os, sys, time, collections
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..')); sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests'))
import numpy as np
from openqp_b200 import basis as B
from openqp_b200.int2 import Int2Compute, Int2RhfData
from openqp_b200.scf import pack
from common import decaying_density
cfg = sys.argv[1]
dump = "/tmp/oqpb_prof.txt"
if os.path.exists(dump): os.unlink(dump)
os.environ["OQPB_PROF_FILE"] = dump
drv = Int2Compute(0)
mol, bs = B.build(cfg)
drv.init(bs); t = time.time(); drv.set_screening(); ts = time.time() - t
d = pack(decaying_density(bs))
drv.run(Int2RhfData(d))
t = time.time(); drv.run(Int2RhfData(d)); tb = time.time() - t; st = drv.last_stats()
drv.profile(True); drv.run(Int2RhfData(d)); tab = drv.profile(False)
tot = sum(v['ms'] for v in tab.values())
print(cfg, bs.describe(), "schwarz %.2fs build %.3fs kernel_ms %.1f quartets %.3e TFLOP/s %.2f (serialised sum %.1f ms)" % (ts, tb, st['kernel_ms'], st['nquartets'], st['flops'] / st['kernel_ms'] / 1e9, tot))
for k, v in sorted(tab.items(), key=lambda kv: -kv[1]['ms']):
    if v['ms'] <= 0: continue
    print("  %-10s %8.2f ms %5.1f%%  q=%.2e prims/q=%7.1f  %6.2f TFLOP/s  %6.1f ns/quartet-SM" % (k, v['ms'], 100 * v['ms'] / tot, v['quartets'], v['prims'] / max(v['quartets'], 1), v['flops'] / v['ms'] / 1e9, v['ms'] * 1e6 * 148 / max(v['quartets'], 1)))
PC = ["ss", "ps", "pp", "ds", "dp", "dd", "fs", "fp", "fd", "ff"]
agg = collections.defaultdict(lambda: [0.0, 0, 0, 0])
bk = collections.defaultdict(lambda: [0.0, 0, 0])
for line in open(dump):
    a, b, ms, n, pr = line.split()
    a, b, ms, n, pr = int(a), int(b), float(ms), int(n), int(pr)
    key = "(%s|%s) b%d%d" % (PC[a // 4], PC[b // 4], a % 4, b % 4)
    g = agg[key]; g[0] += ms; g[1] += n; g[2] += pr; g[3] += 1
    h = bk["b%d%d" % (a % 4, b % 4)]; h[0] += ms; h[1] += n; h[2] += pr
print("by bucket pair (all classes):")
for k, g in sorted(bk.items(), key=lambda kv: -kv[1][0]):
    print("  %s %8.1f ms  q=%.2e prims/q=%.1f" % (k, g[0], g[1], g[2] / max(g[1], 1)))
print("top (class, bucket pair) launches:")
for k, g in sorted(agg.items(), key=lambda kv: -kv[1][0])[:70]:
    print("  %-14s %8.2f ms  launches=%3d q=%.2e prims/q=%7.1f  %7.1f ns/quartet-SM  %6.2f ns/prim-SM" % (k, g[0], g[3], g[1], g[2] / max(g[1], 1), g[0] * 1e6 * 148 / max(g[1], 1), g[0] * 1e6 * 148 / max(g[2], 1)))
