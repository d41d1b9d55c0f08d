import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from openqp_b200 import basis as B
from openqp_b200.int2 import Int2Compute, Int2RhfData
from openqp_b200.scf import pack
from common import decaying_density
drv = Int2Compute(0)
for cfg in sys.argv[1:]:
    mol, bs = B.build(cfg)
    drv.init(bs); t=time.time(); drv.set_screening(); ts=time.time()-t
    d = pack(decaying_density(bs))
    drv.run(Int2RhfData(d))
    t=time.time(); drv.run(Int2RhfData(d)); tb=time.time()-t; st=drv.last_stats()
    drv.profile(True); drv.run(Int2RhfData(d)); tab = drv.profile(False)
    tot = sum(v['ms'] for v in tab.values())
    print(cfg, bs.describe(), "schwarz %.2fs build %.3fs kernel_ms %.1f quartets %.3e TFLOP/s %.2f"%(ts, tb, st['kernel_ms'], st['nquartets'], st['flops']/st['kernel_ms']/1e9))
    for k,v in sorted(tab.items(), key=lambda kv:-kv[1]['ms'])[:25]:
        print("  %-10s %8.2f ms %5.1f%%  q=%.2e prims/q=%7.1f  %6.2f TFLOP/s  %6.1f ns/quartet-SM"%(k, v['ms'], 100*v['ms']/tot, v['quartets'], v['prims']/max(v['quartets'],1), v['flops']/v['ms']/1e9, v['ms']*1e6*148/v['quartets']))
