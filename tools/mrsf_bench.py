#!/usr/bin/env python3
"""MRSF batched multi-density J/K (config 5): python tools/mrsf_bench.py [workload=c5] [nvec=12] [reps=2] [profile=0]
Times oqpb_jk_mrsf (host buffers: H2D of d3, build, D2H of f3) and the device part (kernel_ms)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import numpy as np
from openqp_b200 import basis as B
from openqp_b200 import workloads as W
from openqp_b200.int2 import Int2Compute, Int2MrsfData
cfg = sys.argv[1] if len(sys.argv) > 1 else "c5"
nvec = int(sys.argv[2]) if len(sys.argv) > 2 else 12
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
prof = int(sys.argv[4]) if len(sys.argv) > 4 else 0
mol, bs = B.build(cfg)
drv = Int2Compute(0).init(bs); drv.set_screening()
d3 = W.mrsf_densities(bs, nvec)
for r in range(reps):
    t = time.perf_counter(); c = drv.run(Int2MrsfData(d3, 0.5, 1.0)); dt = time.perf_counter() - t
    st = drv.last_stats()
    print(f"{cfg} {bs.describe()} nvec={nvec} x7: call {dt:.3f}s kernel_ms {st['kernel_ms']:.1f} quartets {st['nquartets']:.3e} "
          f"algorithmic TFLOP/s {st['flops'] / st['kernel_ms'] / 1e9:.2f}  |f3|max {np.abs(c.f3).max():.3e}", flush=True)
if prof:
    drv.profile(True); drv.run(Int2MrsfData(d3, 0.5, 1.0)); tab = drv.profile(False)
    tot = sum(v['ms'] for v in tab.values())
    for k, v in sorted(tab.items(), key=lambda kv: -kv[1]['ms']):
        if v['ms'] > 0:
            print("  %-10s %8.2f ms %5.1f%%  q=%.2e  %6.2f TFLOP/s" % (k, v['ms'], 100 * v['ms'] / tot, v['quartets'], v['flops'] / v['ms'] / 1e9))
