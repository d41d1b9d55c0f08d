#!/usr/bin/env python3
"""Per-rank device time of a partitioned Fock build measured on ONE GPU: every rank's slice (oqpb_set_partition(r, N)) is
built in turn, so the scaling efficiency the N-GPU run can reach (T_1 / (N max_r T_r), no collective) and its split into
imbalance (max/mean) and small-launch tails (N mean / T_1) are known without occupying N GPUs.
usage: python tools/partition_bench.py [workload=w32] [N,N,...=2,4,8]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..')); sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests'))
import numpy as np
from openqp_b200 import basis as B
from openqp_b200.int2 import Int2Compute, Int2RhfData
from openqp_b200.scf import pack
from common import decaying_density
cfg = sys.argv[1] if len(sys.argv) > 1 else "w32"
ns = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "2,4,8").split(",")]
drv = Int2Compute(0)
mol, bs = B.build(cfg)
drv.init(bs); drv.set_screening()
d = pack(decaying_density(bs))
def t_rank(r, n):
    drv.set_partition(r, n)
    drv.run(Int2RhfData(d))
    best = 1e30
    for _ in range(2):
        drv.run(Int2RhfData(d)); best = min(best, drv.last_stats()['kernel_ms'])
    return best
t1 = t_rank(0, 1)
print(f"{cfg} {bs.describe()}: 1 rank {t1:.1f} ms")
for n in ns:
    ts = np.array([t_rank(r, n) for r in range(n)])
    print(f"  N={n}: per-rank ms min {ts.min():.1f} mean {ts.mean():.1f} max {ts.max():.1f}; imbalance max/mean {ts.max() / ts.mean():.3f}; "
          f"tail overhead N*mean/T1 {n * ts.mean() / t1:.3f}; efficiency bound T1/(N*max) {t1 / (n * ts.max()):.3f}")
