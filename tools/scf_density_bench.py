#!/usr/bin/env python3
"""Fock-build times on REAL SCF densities of a benchmark molecule (SURVEY 8d inputs (i)-(iii)): the core-guess density,
every iteration's density, the converged density, and the incremental dD = D_n - D_(n-1) builds the reference runs by default
(scf_addons.F90:2036-2041).  RHF, DIIS; the two-electron part through the GPU builder, the one-electron integrals from the
oracle (test infrastructure: this is a profiling tool, not bench.py).

  python tools/scf_density_bench.py <workload> [out.json]        e.g. w32
"""
import json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from openqp_b200 import workloads as W
from openqp_b200.int2 import Int2Compute, Int2RhfData
from openqp_b200.scf import pack, unpack
from oracle.oracle import Oracle

wl = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", f"r02_scf_{wl}.json")
mol, bs = W.build(wl)
nocc = int(sum(mol.Z)) // 2
t0 = time.time()
o = Oracle(bs)
S, T, V = o.int1e()
H = T + V
enuc = mol.nuclear_repulsion()
t_1e = time.time() - t0
drv = Int2Compute(0).init(bs)
drv.set_screening()
s, U = np.linalg.eigh(S)
X = U @ np.diag(s ** -0.5) @ U.T


def diag(F):
    e, C = np.linalg.eigh(X.T @ F @ X)
    return e, X @ C


def build(dm, label):
    t = time.time()
    c = drv.run(Int2RhfData(pack(dm), post=True))
    wall = time.time() - t
    st = drv.last_stats()
    rec = {"label": label, "max_abs_d": float(np.abs(dm).max()), "quartets": st["nquartets"], "kernel_ms": st["kernel_ms"],
           "wall_ms": 1e3 * wall, "tflops": st["flops"] / max(st["kernel_ms"], 1e-9) / 1e9}
    return unpack(c.f[0], bs.nbf), rec


recs = []
_, C = diag(H)
D = 2.0 * C[:, :nocc] @ C[:, :nocc].T
errs, focks = [], []
F2e_prev, D_prev, e_old = None, None, 0.0
for it in range(40):
    F2e, rec = build(D, f"iter {it} full D" + (" (core guess)" if it == 0 else ""))
    rec["iter"] = it
    if D_prev is not None:
        # the incremental build of the same iteration: dD in, F = F_old + F[dD]  (default path of the reference)
        dF, rinc = build(D - D_prev, f"iter {it} incremental dD")
        rinc["iter"] = it
        rinc["max_dev_vs_full"] = float(np.abs(F2e_prev + dF - F2e).max())
        recs.append(rinc)
    F = H + F2e
    e = enuc + 0.5 * np.sum(D * (H + F))
    err = F @ D @ S - S @ D @ F
    rec["energy"] = float(e); rec["diis_err"] = float(np.abs(err).max())
    recs.append(rec)
    print(f"it {it:2d} E = {e:.10f} err = {np.abs(err).max():.2e}  full {rec['kernel_ms']:.1f} ms / {rec['quartets']:.3e} q"
          + (f"   dD {rinc['kernel_ms']:.1f} ms / {rinc['quartets']:.3e} q (|dD| {rinc['max_abs_d']:.1e})" if D_prev is not None else ""), flush=True)
    if abs(e - e_old) < 1e-9 and np.abs(err).max() < 1e-6:
        break
    e_old = e
    errs.append(err.ravel()); focks.append(F)
    errs, focks = errs[-8:], focks[-8:]
    if len(errs) > 1:
        n = len(errs)
        Bm = -np.ones((n + 1, n + 1)); Bm[n, n] = 0
        for a in range(n):
            for b in range(n):
                Bm[a, b] = errs[a] @ errs[b]
        rhs = np.zeros(n + 1); rhs[n] = -1
        try:
            cc = np.linalg.solve(Bm, rhs)[:n]
            Fd = sum(cc[a] * focks[a] for a in range(n))
        except np.linalg.LinAlgError:
            Fd = F
    else:
        Fd = F
    F2e_prev, D_prev = F2e, D
    _, C = diag(Fd)
    D = 2.0 * C[:, :nocc] @ C[:, :nocc].T
res = {"workload": W.WORKLOADS.get(wl, wl), "nbf": bs.nbf, "nocc": nocc, "energy": float(e), "iterations": it + 1, "one_electron_s": t_1e,
       "peak_tflops": drv.fp64_peak_tflops(), "builds": recs,
       "note": "kernel_ms = CUDA-event time of the build's ERI launches; synthetic-density headline of bench.py for comparison"}
json.dump(res, open(out, "w"), indent=1)
print("wrote", out)
