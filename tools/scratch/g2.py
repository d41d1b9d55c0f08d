import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from openqp_b200 import basis as B
from openqp_b200.int2 import Int2Compute, Int2RhfData
from openqp_b200.scf import pack
sys.path.insert(0, '/root/repo/tests')
from common import decaying_density
drv = Int2Compute(0)
print("fp64 peak", drv.fp64_peak_tflops())
for cfg in sys.argv[1:]:
    mol, bs = B.build(cfg)
    t = time.time(); drv.init(bs); t1 = time.time(); drv.set_screening(); t2 = time.time()
    d = pack(decaying_density(bs))
    for rep in range(2):
        t3 = time.time(); c = drv.run(Int2RhfData(d)); t4 = time.time()
        st = drv.last_stats()
        print(cfg, bs.describe(), "init %.2fs schwarz %.2fs build %.3fs kernel %.1f ms quartets %d skipped %d launches %d GFLOP %.1f -> %.2f TFLOP/s" % (
            t1-t, t2-t1, t4-t3, st['kernel_ms'], st['nquartets'], st['nschwz'], st['launches'], st['flops']/1e9, st['flops']/st['kernel_ms']/1e9), flush=True)
