import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from openqp_b200 import basis as B
from openqp_b200.int2 import Int2Compute, Int2RhfData
from openqp_b200.scf import pack
from common import decaying_density
drv = Int2Compute(0)
mol, bs = B.build(sys.argv[1])
drv.init(bs); drv.set_screening()
d = pack(decaying_density(bs))
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
    drv.run(Int2RhfData(d))
