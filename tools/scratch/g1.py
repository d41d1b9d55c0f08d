import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from openqp_b200 import basis as B
from openqp_b200.int2 import Int2Compute, Int2RhfData, fock_jk
from openqp_b200.scf import pack, unpack
from oracle.oracle import Oracle, rys
drv = Int2Compute(0)
# rys tables
for R in (1,2,3,5,7):
    x = np.array([0.0, 0.3, 1.7, 9.9, 33.3, 38.99, 39.0, 50.0, 74.9, 80.0, 200.0])
    t2, w = drv.rys(R, x)
    err = 0
    for i, xx in enumerate(x):
        u, ww = rys(R, xx)
        err = max(err, np.abs(t2[i] - u/(1+u)).max(), np.abs(w[i]/ww - 1).max())
    print("rys R", R, "max err", err)
for name, molf in (("6-31g(d)", B.water), ("cc-pvtz", B.water)):
    mol = molf(); bs = B.BasisSet(mol, name)
    o = Oracle(bs); Qo = o.set_screening()
    drv.init(bs); t=time.time(); Qg = drv.set_screening(); print(name, "schwarz time", time.time()-t)
    print(name, "schwarz max rel diff", np.abs(Qg-Qo).max(), np.abs(Qg/Qo-1).max())
    # blocks
    rng = np.random.default_rng(0)
    worst = 0
    for it in range(60):
        i,j,k,l = rng.integers(0, bs.nshell, 4)
        bo = o.eri_block(i,j,k,l); bg = drv.eri_block(i,j,k,l)
        if bo.shape != bg.shape: print("shape mismatch", i,j,k,l, bo.shape, bg.shape); continue
        e = np.abs(bo-bg).max()
        if e > 1e-11: print("block", i,j,k,l, bs.am[[i,j,k,l]], e, np.abs(bo).max())
        worst = max(worst, e)
    print(name, "worst block err", worst)
    Dm = rng.normal(size=(bs.nbf,bs.nbf)); Dm = Dm+Dm.T
    fo, so = o.fock(pack(Dm))
    fg, ns = fock_jk(drv, pack(Dm))
    print(name, "fock err", np.abs(fo-fg).max(), "nschwz", so['nschwz'], ns, drv.last_stats(), so)
print("fp64 peak", drv.fp64_peak_tflops())
