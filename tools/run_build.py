#!/usr/bin/env python3
"""usage: python tools/run_build.py <workload> [nbuilds]  -- N RHF Fock builds of a workload (target of the ncu helper scripts)"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..')); sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests'))
from openqp_b200 import basis as B
from openqp_b200.int2 import Int2Compute, Int2RhfData
from openqp_b200.scf import pack
from common import decaying_density
cfg = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
drv = Int2Compute(0)
mol, bs = B.build(cfg)
drv.init(bs); drv.set_screening()
d = pack(decaying_density(bs))
for _ in range(n):
    t = time.time(); drv.run(Int2RhfData(d)); tb = time.time() - t
st = drv.last_stats()
print(cfg, bs.describe(), "build %.3fs kernel_ms %.1f quartets %.3e" % (tb, st['kernel_ms'], st['nquartets']))
