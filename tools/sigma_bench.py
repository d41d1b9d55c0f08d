#!/usr/bin/env python3
"""MRSF sigma session (routec_sig_iter) on the config-5 molecule: wall time of one Davidson step for nv trial vectors through
the device session, next to the same step with the J/K build alone through host buffers (oqpb_jk_mrsf: H2D of the 7 nv
densities, D2H of their images -- what a host-side sigma triple would move).
usage: python tools/sigma_bench.py [workload=c5] [nv=12] [cutoff=1e-8]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import numpy as np
from openqp_b200 import workloads as W
from openqp_b200.int2 import Int2Compute, Int2MrsfData, RoutecSig
cfg = sys.argv[1] if len(sys.argv) > 1 else "c5"
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 12
cutoff = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-8
mol, bs = W.build(cfg)
drv = Int2Compute(0).init(bs, cutoff); drv.set_screening()
n = bs.nbf
rng = np.random.default_rng(3)
Cm = np.linalg.qr(rng.normal(size=(n, n)))[0]
fa = np.diag(np.linspace(-10, 3, n)); fb = fa.copy()
nel = int(sum(mol.Z)); na, nb = nel // 2 + 1, nel // 2 - 1
ntrial = na * (n - nb)
X = rng.normal(size=(ntrial, nv)) * np.exp(-rng.uniform(0, 6, size=(ntrial, 1)))
sig = RoutecSig(drv)
assert sig.begin(Cm, Cm, fa, fb, na, nb, 1, 0.5) == 0
for r in range(3):
    t = time.perf_counter(); s = sig.apply(X); dt = time.perf_counter() - t
    st = drv.last_stats()
    print(f"{cfg} nbf {n} ntrial {ntrial} nv {nv} cutoff {cutoff:g}: routec_sig_iter {1e3 * dt:.1f} ms (J/K kernels {st['kernel_ms']:.1f} ms, "
          f"{st['nquartets']:.3e} quartets); bytes over the bus {2 * X.nbytes / 1e6:.1f} MB", flush=True)
sig.end()
d3 = W.mrsf_densities(bs, nv)
for r in range(2):
    t = time.perf_counter(); c = drv.run(Int2MrsfData(d3, 0.5, 0.5)); dt = time.perf_counter() - t
    print(f"   oqpb_jk_mrsf with host buffers (synthetic densities): {1e3 * dt:.1f} ms (kernels {drv.last_stats()['kernel_ms']:.1f} ms); "
          f"bytes over the bus {2 * d3.nbytes / 1e6:.1f} MB", flush=True)
