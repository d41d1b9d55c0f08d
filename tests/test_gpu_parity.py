"""GPU: parity of the CUDA path (through the C ABI) against the oracle, the committed golden vectors and the
reference's golden energies.  Tolerances: integer/list work bit-exact; Fock elements 1e-10 Eh absolute;
SCF energies 1e-8 Eh (BASELINE.json north_star)."""
import ctypes

import numpy as np
import pytest

from common import decaying_density, golden, random_sym_density, rpa_energies
from openqp_b200 import basis as B
from openqp_b200.scf import pack, scf, unpack

pytestmark = pytest.mark.gpu

FOCK_TOL = 1e-10


@pytest.fixture(scope="module")
def drv():
    from openqp_b200.int2 import Int2Compute
    d = Int2Compute(0)
    yield d
    d.clean()


def _pair(oracle_mod, drv, mol, basis, cutoff=5e-11, upload_q=False):
    bs = B.BasisSet(mol, basis)
    o = oracle_mod.Oracle(bs, cutoff)
    q = o.set_screening()
    drv.init(bs, cutoff)
    drv.set_screening(q if upload_q else None)
    return bs, o


def test_rys_tables_vs_oracle(oracle_mod, drv):
    for R in range(1, 8):
        x = np.concatenate([np.linspace(0, 80, 161), [1e-9, 38.999, 39.0, 74.999, 75.0, 150.0, 1e3]])
        t2, w = drv.rys(R, x)
        for i, xx in enumerate(x):
            u, ww = oracle_mod.rys(R, xx)
            o = np.argsort(u)
            assert np.abs(t2[i] - (u / (1 + u))[o]).max() < 5e-13, (R, xx)
            assert np.abs(w[i] / ww[o] - 1).max() < 5e-12, (R, xx)


@pytest.mark.parametrize("basis", ["6-31g(d)", "cc-pvdz", "cc-pvtz"])
def test_schwarz_matrix(oracle_mod, drv, basis):
    """ints_exchange (int2.F90:1582-1737) on the device."""
    bs, o = _pair(oracle_mod, drv, B.water_dimer(), basis)
    qg, qo = drv.schwarz(), o.schwarz
    assert np.abs(qg - qo).max() < 1e-12
    m = qo > 1e-10
    assert np.abs(qg[m] / qo[m] - 1).max() < 1e-10


@pytest.mark.parametrize("basis", ["6-31g(d)", "cc-pvtz"])
def test_eri_blocks(oracle_mod, drv, basis):
    """shellquartet (int2.F90:1051-1183): every angular-momentum class, all index orders."""
    bs, o = _pair(oracle_mod, drv, B.water(), basis)
    rng = np.random.default_rng(0)
    seen = set()
    for it in range(400):
        i, j, k, l = (int(x) for x in rng.integers(0, bs.nshell, 4))
        key = tuple(sorted((tuple(sorted((bs.am[i], bs.am[j]))), tuple(sorted((bs.am[k], bs.am[l]))))))
        if key in seen and it > 150:
            continue
        seen.add(key)
        ci, cj, ck, cl = max(i, j), min(i, j), max(k, l), min(k, l)
        bo = o.eri_block(ci, cj, ck, cl)  # the reference engine is only ever called with i>=j, k>=l
        if i < j:
            bo = bo.transpose(1, 0, 2, 3)
        if k < l:
            bo = bo.transpose(0, 1, 3, 2)
        bg = drv.eri_block(i, j, k, l)
        assert bg.shape == bo.shape
        assert np.abs(bg - bo).max() < 2e-12, (i, j, k, l, bs.am[[i, j, k, l]])
    assert len(seen) >= (15 if basis == "6-31g(d)" else 40)


@pytest.mark.parametrize("basis", ["6-31g(d)", "cc-pvtz"])
def test_fock_vs_committed_golden(drv, basis):
    """no live oracle: compare against tests/golden/oracle_fock_h2o.json"""
    from openqp_b200.int2 import fock_jk
    g = golden("oracle_fock_h2o.json")[basis]
    bs = B.BasisSet(B.water(), basis)
    drv.init(bs)
    q = drv.set_screening()
    assert abs(q.sum() - g["schwarz_sum"]) < 1e-9
    d = random_sym_density(bs.nbf, g["seed"])
    f, nschwz = fock_jk(drv, pack(d))
    assert np.abs(f[0] - np.array(g["fock"])).max() < FOCK_TOL
    assert nschwz == g["stats"]["nschwz"]
    assert drv.last_stats()["nquartets"] == g["stats"]["nquartets"]


@pytest.mark.parametrize("molname,basis", [("benzene", "cc-pvdz"), ("dimer", "cc-pvtz"), ("benzene", "6-31g(d)")])
def test_fock_rhf_and_screening_bit_exact(oracle_mod, drv, molname, basis):
    """int2_twoei + int2_rhf_data_t (int2.F90:589-923, 1414-1484) with a decaying density so that screening bites.
    With the oracle's Schwarz matrix uploaded the surviving quartet set must be identical."""
    from openqp_b200.int2 import Int2RhfData
    mol = B.benzene() if molname == "benzene" else B.water_dimer()
    bs, o = _pair(oracle_mod, drv, mol, basis, upload_q=True)
    dm = decaying_density(bs) * (1e-4 if molname == "benzene" else 1e-2)
    d = pack(dm)
    lst, n, nschwz = o.quartet_list(d)
    drv.record_quartets(True)
    cons = drv.run(Int2RhfData(d, scale_exchange=0.2, scale_coulomb=1.0))  # B3LYP-like hybrid (config 2)
    drv.record_quartets(False)
    assert cons.skipped == nschwz and nschwz > 0
    got = drv.quartets()
    assert got.shape == lst.shape
    key = lambda a: a[np.lexsort(a.T[::-1])]
    assert np.array_equal(key(got), key(lst))  # bit-exact surviving quartet list
    dsh_g, md_g = drv.shell_density()
    dsh_o, md_o = o.shlden(0, np.atleast_2d(d), 1)
    assert np.array_equal(dsh_g, dsh_o) and md_g == md_o  # shlden (int2.F90:999-1047) bit-exact
    f, st = o.fock(d, scale_exchange=0.2, scale_coulomb=1.0, post=False)
    assert np.abs(cons.f - f).max() < FOCK_TOL * max(1.0, 0.0)
    # same build with the device-computed Schwarz matrix: counts may differ only through ulp-level Q differences
    drv.set_screening(None)
    cons2 = drv.run(Int2RhfData(d, scale_exchange=0.2, scale_coulomb=1.0))
    assert abs(cons2.skipped - nschwz) <= max(2, nschwz // 100000)
    assert np.abs(cons2.f - f).max() < FOCK_TOL


def test_fock_urohf(oracle_mod, drv):
    """int2_urohf_data_t (int2.F90:1488-1578)."""
    from openqp_b200.int2 import fock_jk
    bs, o = _pair(oracle_mod, drv, B.water_dimer(), "cc-pvdz")
    da, db = pack(random_sym_density(bs.nbf, 21)), pack(random_sym_density(bs.nbf, 22))
    d = np.stack([da, db])
    f, _ = fock_jk(drv, d, 0.5, 0.8, urohf=True)
    fo, _ = o.fock(d, 0.5, 0.8, urohf=True)
    assert np.abs(f - fo).max() < 5e-10  # |D| ~ 3: absolute tolerance scaled


def test_multi_fock_rhf(oracle_mod, drv):
    """nfocks > 1 closed-shell builds (CPHF / Hessian callers, modules/cphf.F90:611)."""
    from openqp_b200.int2 import fock_jk
    bs, o = _pair(oracle_mod, drv, B.water(), "cc-pvtz")
    d = np.stack([pack(random_sym_density(bs.nbf, s, 0.1)) for s in (1, 2, 3)])
    f, _ = fock_jk(drv, d)
    fo, _ = o.fock(d)
    assert np.abs(f - fo).max() < FOCK_TOL


def test_scf_energy_golden_through_gpu(oracle_mod, drv):
    """Config 1: RHF/6-31G(d) water, energy vs examples/HF/H2O_RHF-HF_ENERGY.json within 1e-8 Eh;
    1e integrals come from the oracle (out of scope of the builder)."""
    from openqp_b200.int2 import fock_jk
    ref = golden("reference_energies.json")["h2o_rhf_631gd"]["energy"]
    mol = B.water()
    bs, o = _pair(oracle_mod, drv, mol, "6-31g(d)")
    S, T, V = o.int1e()
    e, D, F = scf(bs.nbf, S, T + V, mol.nuclear_repulsion(), lambda dp: fock_jk(drv, dp)[0], 5)
    assert abs(e - ref) < 1e-8
    e_o, _, _ = scf(bs.nbf, S, T + V, mol.nuclear_repulsion(), lambda dp: o.fock(dp)[0], 5)
    assert abs(e - e_o) < 1e-10
    ref_u = golden("reference_energies.json")["h2o_uhf_triplet_631gd"]["energy"]
    e_u, _, _ = scf(bs.nbf, S, T + V, mol.nuclear_repulsion(), lambda dp: fock_jk(drv, dp, urohf=True)[0], 6, 4)
    assert abs(e_u - ref_u) < 1e-8


def test_td_consumer(oracle_mod, drv):
    """int2_td_data_t (tdhf_lib.F90:140-224): A+B, A-B and TDA variants."""
    from openqp_b200.int2 import Int2TdData
    bs, o = _pair(oracle_mod, drv, B.water(), "cc-pvdz", cutoff=1e-8)
    rng = np.random.default_rng(5)
    P = rng.normal(size=(3, bs.nbf, bs.nbf)) * 0.1
    c = drv.run(Int2TdData(P, int_apb=True, int_amb=True, scale_exchange=0.5))
    apb, amb, st = o.td(P, 0.5, 1.0, int_apb=True, int_amb=True)
    assert np.abs(c.apb - apb).max() < 1e-10 and np.abs(c.amb - amb).max() < 1e-10
    assert c.skipped == st["nschwz"]
    c = drv.run(Int2TdData(P, tamm_dancoff=True, tamm_dancoff_coulomb=True))
    apb, amb, st = o.td(P, tamm_dancoff=True, tamm_dancoff_coulomb=True)
    assert np.abs(c.amb - amb).max() < 1e-10
    c = drv.run(Int2TdData(P, tamm_dancoff=True))
    apb, amb, st = o.td(P, tamm_dancoff=True)
    assert np.abs(c.amb - amb).max() < 1e-10


def test_tdhf_golden_through_gpu(oracle_mod, drv):
    from openqp_b200.int2 import Int2TdData, fock_jk
    g = golden("reference_energies.json")["h2o_tdhf_631gd"]
    mol = B.water()
    bs, o = _pair(oracle_mod, drv, mol, "6-31g(d)")
    S, T, V = o.int1e()

    def td(P):
        c = drv.run(Int2TdData(P, int_apb=True, int_amb=True))
        return c.apb, c.amb

    e, w = rpa_energies(bs, S, T + V, mol.nuclear_repulsion(), 5, lambda dp: fock_jk(drv, dp)[0], td)
    assert abs(e - g["energy"]) < 1e-8
    assert np.allclose(w, g["td_energies"], atol=2e-7)


def test_mrsf_consumer(oracle_mod, drv):
    """int2_mrsf_data_t (tdhf_mrsf_lib.F90:218-333): nvec x 7 batched densities (config 5 shape, small)."""
    from openqp_b200.int2 import Int2MrsfData
    bs, o = _pair(oracle_mod, drv, B.water_dimer(), "6-31g(d)", cutoff=1e-8)
    rng = np.random.default_rng(7)
    d3 = rng.normal(size=(3, 7, bs.nbf, bs.nbf)) * 0.1
    c = drv.run(Int2MrsfData(d3, scale_exchange=0.5, scale_coulomb=0.5))
    f3, st = o.mrsf(d3, 0.5, 0.5)
    assert np.abs(c.f3 - f3).max() < 1e-10
    assert c.skipped == st["nschwz"]


def test_mrsf_consumer_dmma_pure_d(oracle_mod, drv):
    """Batched multi-density digestion on the FP64 tensor cores (DMMA m8n8k4, eri_kernel.cuh gen_contract): pure 5d
    shells, a matrix count (5 x 7 = 35) that is not a multiple of the 8-row MMA tile, Coulomb on 20 of the 35."""
    from openqp_b200.int2 import Int2MrsfData
    bs, o = _pair(oracle_mod, drv, B.water(), "cc-pvdz", cutoff=1e-9)
    rng = np.random.default_rng(11)
    d3 = rng.normal(size=(5, 7, bs.nbf, bs.nbf)) * 0.1
    c = drv.run(Int2MrsfData(d3, scale_exchange=0.7, scale_coulomb=0.9))
    f3, st = o.mrsf(d3, 0.7, 0.9)
    assert np.abs(c.f3 - f3).max() < 1e-10
    assert c.skipped == st["nschwz"]


def test_legacy_seam_routec_fock_jk(oracle_mod, drv):
    """routec_fock_jk (routec_bridge.F90:33-40): by-reference scalars, f returned ready to use, info = 0."""
    from openqp_b200.int2 import lib
    bs, o = _pair(oracle_mod, drv, B.water(), "6-31g(d)")
    d = np.ascontiguousarray(pack(random_sym_density(bs.nbf, 3)))
    f = np.zeros_like(d)
    L = lib()
    L.oqpb_set_default_ctx(drv._h)
    info, nbf, nf = ctypes.c_int(7), ctypes.c_int(bs.nbf), ctypes.c_int(1)
    se, sc = ctypes.c_double(1.0), ctypes.c_double(1.0)
    L.routec_fock_jk(d.ctypes.data_as(ctypes.c_void_p), f.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nbf),
                     ctypes.byref(nf), ctypes.byref(se), ctypes.byref(sc), ctypes.byref(info))
    assert info.value == 0
    fo, _ = o.fock(d)
    assert np.abs(f - fo[0]).max() < FOCK_TOL
    nbf_bad = ctypes.c_int(bs.nbf + 1)
    L.routec_fock_jk(d.ctypes.data_as(ctypes.c_void_p), f.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nbf_bad),
                     ctypes.byref(nf), ctypes.byref(se), ctypes.byref(sc), ctypes.byref(info))
    assert info.value != 0  # declines -> native fallback on the Fortran side
    L.oqpb_set_default_ctx(None)


def test_partition_sums_to_full(drv):
    """Replicated-data split (int2.F90:759-761): partial Focks of 3 ranks add up to the full build."""
    from openqp_b200.int2 import Int2RhfData
    bs = B.BasisSet(B.water_dimer(), "cc-pvdz")
    drv.init(bs)
    drv.set_screening()
    d = pack(decaying_density(bs))
    full = drv.run(Int2RhfData(d))
    nq = drv.last_stats()["nquartets"]
    acc, nqs, skipped = np.zeros_like(full.f), 0, 0
    for r in range(3):
        drv.set_partition(r, 3)
        c = drv.run(Int2RhfData(d))
        acc += c.f
        nqs += drv.last_stats()["nquartets"]
        skipped += c.skipped
    drv.set_partition(0, 1)
    assert nqs == nq and skipped == full.skipped
    assert np.abs(acc - full.f).max() < 1e-11


def test_full_size_properties_c20h42(drv):
    """Config 3 (n-C20H42/def2-SVP, 490 bf) -- too large for the serial oracle in seconds, so check
    size-independent properties: linearity of the build, J/K scale separation, and agreement between the
    SYM consumer and the GEN (TD, Tamm-Dancoff+Coulomb) consumer on the same symmetric density."""
    from openqp_b200.int2 import Int2RhfData, Int2TdData
    mol, bs = B.build("c3")
    drv.init(bs, 1e-12)
    drv.set_screening()
    d1, d2 = decaying_density(bs, 1), decaying_density(bs, 2)
    f1 = drv.run(Int2RhfData(pack(d1), post=True)).f[0]
    f2 = drv.run(Int2RhfData(pack(d2), post=True)).f[0]
    f12 = drv.run(Int2RhfData(pack(d1 + 2 * d2), post=True)).f[0]
    assert np.abs(f12 - (f1 + 2 * f2)).max() < 1e-9
    fj = drv.run(Int2RhfData(pack(d1), scale_exchange=0.0, post=True)).f[0]
    fk = drv.run(Int2RhfData(pack(d1), scale_coulomb=0.0, post=True)).f[0]
    assert np.abs(fj + fk - f1).max() < 1e-9
    c = drv.run(Int2TdData(d1[None], tamm_dancoff=True, tamm_dancoff_coulomb=True))
    # TDA amb = J[P+P^T] - K[P]  (= 2J - K for symmetric P) ;  RHF f = J - K/2
    assert np.abs(c.amb[0] - 2 * unpack(f1, bs.nbf)).max() < 2e-9


def test_mrsf_device_entry_matches_host_entry(drv):
    """oqpb_jk_mrsf_dev (device-resident d3 / f3) against oqpb_jk_mrsf (host buffers) on the same input."""
    import torch
    from openqp_b200.int2 import Int2MrsfData
    bs = B.BasisSet(B.water_dimer(), "6-31g(d)")
    drv.init(bs)
    drv.set_screening()
    rng = np.random.default_rng(5)
    d3 = rng.normal(size=(2, 7, bs.nbf, bs.nbf)) * 0.1
    host = drv.run(Int2MrsfData(d3, scale_exchange=0.5, scale_coulomb=1.0)).f3
    dF = np.ascontiguousarray(np.transpose(d3, (3, 2, 1, 0)))
    d_dev = torch.from_numpy(dF).cuda()
    f_dev = torch.full_like(d_dev, 7.0)  # must be overwritten, not accumulated into
    drv.mrsf_dev(d_dev.data_ptr(), f_dev.data_ptr(), 2, 7, scale_exchange=0.5, scale_coulomb=1.0)
    drv.synchronize()
    torch.cuda.synchronize()
    dev = np.transpose(f_dev.cpu().numpy(), (3, 2, 1, 0))
    assert np.abs(dev - host).max() < 1e-12


def test_full_size_properties_w32_and_mrsf(drv):
    """Headline workload ((H2O)32 cc-pVTZ 5d/7f, 1856 bf: every s..f class, group and team kernels) and the batched
    multi-density consumer at that size: linearity of the SYM build, and MRSF components of a symmetric density against
    the RHF build of the same density (c <= 4: J - K = F_rhf with scale_exchange 2;  c > 4: -K)."""
    from openqp_b200.int2 import Int2MrsfData, Int2RhfData
    mol, bs = B.build("w32")
    drv.init(bs)
    drv.set_screening()
    d1, d2 = decaying_density(bs, 1), decaying_density(bs, 2)
    f1 = drv.run(Int2RhfData(pack(d1), post=True)).f[0]
    f2 = drv.run(Int2RhfData(pack(d2), post=True)).f[0]
    f12 = drv.run(Int2RhfData(pack(d1 - 0.5 * d2), post=True)).f[0]
    assert np.abs(f12 - (f1 - 0.5 * f2)).max() < 1e-9
    fk = drv.run(Int2RhfData(pack(d1), scale_coulomb=0.0, post=True)).f[0]
    mol5, bs5 = B.build("c3")
    drv.init(bs5, 1e-12)
    drv.set_screening()
    d = decaying_density(bs5, 3)
    fr = drv.run(Int2RhfData(pack(d), scale_exchange=2.0, post=True)).f[0]  # J - K
    frk = drv.run(Int2RhfData(pack(d), scale_coulomb=0.0, post=True)).f[0]  # -K/2
    d3 = np.broadcast_to(d, (1, 7, bs5.nbf, bs5.nbf)).copy()
    f3 = drv.run(Int2MrsfData(d3, scale_exchange=1.0, scale_coulomb=1.0)).f3
    assert np.abs(f3[0, 0] - unpack(fr, bs5.nbf)).max() < 2e-9
    assert np.abs(f3[0, 3] - f3[0, 0]).max() < 1e-10
    assert np.abs(f3[0, 6] - 2 * unpack(frk, bs5.nbf)).max() < 2e-9
    assert np.isfinite(fk).all()


def test_cam_two_pass_build(oracle_mod, drv):
    """Range-separated build int2_run_cam (int2.F90:538-584): regular pass + Erf-attenuated pass (int_rys.F90:225-227)
    screened with the attenuated Schwarz matrix (int2.F90:674-685), s..f shells, RHF and UROHF consumers."""
    from openqp_b200.int2 import Int2RhfData, Int2UrohfData
    bs, o = _pair(oracle_mod, drv, B.water_dimer(), "cc-pvtz", upload_q=True)
    mu, alpha, beta = 0.33, 0.19, 0.46  # CAM-B3LYP
    qa_o = o.schwarz_attenuated(mu)
    qa_g = drv.set_screening_cam(mu)
    assert np.abs(qa_g - qa_o).max() <= 1e-10 * max(1.0, qa_o.max())
    drv.set_screening_cam(mu, qa_o)  # the oracle's bounds: the screening decisions must then be identical
    d = pack(decaying_density(bs) * 1e-2)
    c = drv.run(Int2RhfData(d, post=True), cam=True, alpha=alpha, beta=beta, mu=mu)
    fo, st = o.fock_cam(d, alpha, beta, mu)
    assert np.abs(c.f - fo).max() < FOCK_TOL
    assert c.skipped == st["nschwz"] and st["nschwz"] > 0
    # the attenuated pass alone (alpha = 0, no regular Coulomb): exercises the pass-2 integrals in isolation
    c2 = drv.run(Int2RhfData(d, post=True), cam=True, alpha=0.0, beta=1.0, mu=mu, alpha_coulomb=0.0, beta_coulomb=1.0)
    fo2, _ = o.fock_cam(d, 0.0, 1.0, mu, alpha_coulomb=0.0, beta_coulomb=1.0)
    assert np.abs(c2.f - fo2).max() < FOCK_TOL and np.abs(fo2).max() > 1e-3
    da, db = pack(random_sym_density(bs.nbf, 31, 0.05)), pack(random_sym_density(bs.nbf, 32, 0.05))
    du = np.stack([da, db])
    cu = drv.run(Int2UrohfData(du, post=True), cam=True, alpha=alpha, beta=beta, mu=mu)
    fu, _ = o.fock_cam(du, alpha, beta, mu, urohf=True)
    # dense (non-decaying) random density, |D| up to ~0.2, |F| ~ 0.5: nothing is screened, so every integral within rounding
    # of the element cutoff 5e-11 (int2.F90:1806-1812) that falls on the other side than in the oracle moves a Fock element by
    # up to cutoff * |D| * (a few); the sum over those flips is 2.0-2.2e-10 here and depends on the kernel's summation order
    # (measured 2.10e-10 / 2.16e-10 with different kernel selections).  The screened, decaying densities above hold 1e-10.
    assert np.abs(cu.f - fu).max() < 4 * FOCK_TOL
    # a regular build afterwards is unaffected by the cached attenuated data
    c3 = drv.run(Int2RhfData(d, post=True))
    f3, _ = o.fock(d)
    assert np.abs(c3.f - f3).max() < FOCK_TOL


def test_cam_response_consumers(oracle_mod, drv):
    """int2_run_cam with the TD consumer (same update in both passes) and the MRSF consumer (pass 2 = attenuated
    exchange of component 7 only, tdhf_mrsf_lib.F90:312-326)."""
    from openqp_b200.int2 import Int2MrsfData, Int2TdData
    bs, o = _pair(oracle_mod, drv, B.water_dimer(), "6-31g(d)", upload_q=True)
    mu, alpha, beta = 0.33, 0.19, 0.46
    drv.set_screening_cam(mu, o.schwarz_attenuated(mu))  # identical bounds on both sides: identical quartet lists
    rng = np.random.default_rng(3)
    d3 = rng.normal(size=(3, 7, bs.nbf, bs.nbf)) * 0.1
    c = drv.run(Int2MrsfData(d3), cam=True, alpha=alpha, beta=beta, mu=mu)
    f3, _ = o.mrsf_cam(d3, alpha, beta, mu)
    assert np.abs(c.f3 - f3).max() < FOCK_TOL
    f3_reg, _ = o.mrsf(d3, alpha, 1.0)
    assert np.abs(f3[:, :6] - f3_reg[:, :6]).max() < 1e-14 and np.abs(f3[:, 6] - f3_reg[:, 6]).max() > 1e-4
    d2 = rng.normal(size=(2, bs.nbf, bs.nbf)) * 0.1
    ct = drv.run(Int2TdData(d2, int_apb=True, int_amb=True), cam=True, alpha=alpha, beta=beta, mu=mu)
    apb, amb, _ = o.td_cam(d2, alpha, beta, mu, int_apb=True, int_amb=True)
    assert np.abs(ct.amb - amb).max() < FOCK_TOL, np.abs(ct.amb - amb).max()
    assert np.abs(ct.apb - apb).max() < FOCK_TOL, np.abs(ct.apb - apb).max()


def test_edge_cases(oracle_mod, drv):
    """Degenerate inputs: a one-shell basis, shell pairs without surviving primitives (centres 60 bohr apart), a zero
    density (every bra pair skipped), the maximum number of Fock matrices, a density count that is not a multiple of any
    tile (MRSF, 10 x 7 = 70 matrices), and a build after a zero build (plan cache keyed on the density bound)."""
    from openqp_b200.int2 import Int2MrsfData, Int2RhfData
    # (1) hydrogen atom, STO-3G: one s shell, one quartet
    h = B.Molecule(np.array([1]), np.zeros((1, 3)), "H")
    bs = B.BasisSet(h, "sto-3g")
    o = oracle_mod.Oracle(bs); q = o.set_screening()
    drv.init(bs); drv.set_screening(q)
    d = np.array([[0.7]])
    c = drv.run(Int2RhfData(d, post=True))
    fo, st = o.fock(d)
    assert bs.nshell == 1 and np.abs(c.f - fo).max() < 1e-13 and c.skipped == st["nschwz"] == 0
    # (2) two water molecules 60 bohr apart: inter-molecular shell pairs lose all their primitives
    w = B.water()
    far = B.Molecule(np.concatenate([w.Z, w.Z]), np.concatenate([w.xyz, w.xyz + np.array([60.0, 0.0, 0.0])]), "far dimer")
    bs = B.BasisSet(far, "cc-pvdz")
    o = oracle_mod.Oracle(bs); q = o.set_screening()
    drv.init(bs); drv.set_screening(q)
    assert (q[:bs.nshell // 2, bs.nshell // 2:] == 0).any()
    dm = pack(random_sym_density(bs.nbf, 5, 0.1))
    # (3) zero density first: everything is skipped at the bra level, F = 0
    z = drv.run(Int2RhfData(np.zeros_like(dm), post=True))
    _, stz = o.fock(np.zeros_like(dm))
    assert np.abs(z.f).max() == 0.0 and z.skipped == stz["nschwz"] and drv.last_stats()["nquartets"] == 0
    c = drv.run(Int2RhfData(dm, post=True))
    fo, st = o.fock(dm)
    assert np.abs(c.f - fo).max() < FOCK_TOL and c.skipped == st["nschwz"]
    # (4) seven Fock matrices in one pass (MAX_MATS - 1)
    ds = np.stack([pack(random_sym_density(bs.nbf, 40 + k, 0.1)) for k in range(7)])
    c7 = drv.run(Int2RhfData(ds, scale_exchange=0.3, post=True))
    f7, _ = o.fock(ds, 0.3)
    assert np.abs(c7.f - f7).max() < FOCK_TOL
    # (5) 70 general densities
    rng = np.random.default_rng(9)
    d3 = rng.normal(size=(10, 7, bs.nbf, bs.nbf)) * 0.05
    cm = drv.run(Int2MrsfData(d3, scale_exchange=0.5, scale_coulomb=1.0))
    fm, stm = o.mrsf(d3, 0.5, 1.0)
    assert np.abs(cm.f3 - fm).max() < FOCK_TOL and cm.skipped == stm["nschwz"]


def test_generic_jk_against_dense_eri(oracle_mod, drv):
    """oqpb_jk: J[P](a,b) = sum (ab|cd) P(c,d), K[P](a,c) = sum (ab|cd) P(b,d) for general (non-symmetric) P, against the
    oracle's dense ERI tensor; J-only / K-only selection leaves the other slabs untouched."""
    from openqp_b200.int2 import jk
    bs, o = _pair(oracle_mod, drv, B.water(), "cc-pvdz", cutoff=1e-14)
    eri = o.dense_eri()
    rng = np.random.default_rng(17)
    P = rng.normal(size=(3, bs.nbf, bs.nbf)) * 0.1
    J, K, _ = jk(drv, P, want_j=[1, 0, 1], want_k=[1, 1, 0])
    for m in range(3):
        Jr = np.einsum("abcd,cd->ab", eri, P[m])
        Kr = np.einsum("abcd,bd->ac", eri, P[m])
        if m != 1:
            assert np.abs(J[m] - Jr).max() < 1e-11
        else:
            assert np.abs(J[m]).max() == 0.0
        if m != 2:
            assert np.abs(K[m] - Kr).max() < 1e-11
        else:
            assert np.abs(K[m]).max() == 0.0


def test_gradient_response_consumers(oracle_mod, drv):
    """int2_tdgrd_data_t (tdhf_lib.F90:228-295), int2_rpagrd_data_t (:1068-1320, nspin 1 and 2) and int2_umrsf_data_t
    (tdhf_mrsf_lib.F90:337-426) through the generic J/K engine, against the oracle's line-by-line restatements
    (quartet lists identical: same Schwarz matrix, same shell densities)."""
    from openqp_b200.int2 import Int2RpagrdData, Int2TdgrdData, Int2UmrsfData
    bs, o = _pair(oracle_mod, drv, B.water_dimer(), "cc-pvdz", cutoff=1e-9, upload_q=True)
    rng = np.random.default_rng(23)
    n = bs.nbf
    se, sc = 0.6, 0.8
    d2 = rng.normal(size=(2, n, n)) * 0.1
    c = drv.run(Int2TdgrdData(d2, int_apb=True, int_amb=True, scale_exchange=se, scale_coulomb=sc))
    apb, amb, st = o.tdgrd(d2, se, sc, True, True)
    assert np.abs(c.apb - apb).max() < 1e-10 and np.abs(c.amb - amb).max() < 1e-10
    assert c.skipped == st["nschwz"] and st["nschwz"] > 0
    c = drv.run(Int2TdgrdData(d2, int_apb=True, int_amb=False, scale_exchange=se, scale_coulomb=sc))
    apb, amb, st = o.tdgrd(d2, se, sc, True, False)
    assert np.abs(c.apb - apb).max() < 1e-10 and np.abs(c.amb).max() == 0.0
    sym = lambda a: a + np.swapaxes(a, -1, -2)
    xpy, t = sym(rng.normal(size=(2, 1, n, n)) * 0.1), sym(rng.normal(size=(1, 1, n, n)) * 0.1)
    xmy = rng.normal(size=(2, 1, n, n)) * 0.1
    c = drv.run(Int2RpagrdData(xpy, xmy, t, 1, se, sc))
    hpp, hpt, hmm, st = o.rpagrd(xpy, xmy, t, 1, se, sc)
    assert np.abs(c.hpp - hpp).max() < 1e-10 and np.abs(c.hpt - hpt).max() < 1e-10 and np.abs(c.hmm - hmm).max() < 1e-10
    assert c.skipped == st["nschwz"]
    c = drv.run(Int2RpagrdData(None, None, t, 1, se, sc))  # the H+[T+Z] call of tdhf_z_vector.F90:341
    _, hpt, _, st = o.rpagrd(None, None, t, 1, se, sc)
    assert np.abs(c.hpt - hpt).max() < 1e-10 and c.skipped == st["nschwz"]
    xpy2, xmy2 = rng.normal(size=(1, 2, n, n)) * 0.1, rng.normal(size=(1, 2, n, n)) * 0.1
    c = drv.run(Int2RpagrdData(xpy2, xmy2, None, 2, se, sc))
    hpp, _, hmm, st = o.rpagrd(xpy2, xmy2, None, 2, se, sc)
    assert np.abs(c.hpp - hpp).max() < 1e-10 and np.abs(c.hmm - hmm).max() < 1e-10 and c.skipped == st["nschwz"]
    d3 = rng.normal(size=(2, 11, n, n)) * 0.1
    c = drv.run(Int2UmrsfData(d3, se, sc))
    f3, st = o.umrsf(d3, se, sc)
    assert np.abs(c.f3 - f3).max() < 1e-10 and c.skipped == st["nschwz"]
    # range-separated variant: pass 2 = attenuated exchange of component 11 only
    mu, alpha, beta = 0.33, 0.19, 0.46
    drv.set_screening_cam(mu, o.schwarz_attenuated(mu))
    c = drv.run(Int2UmrsfData(d3), cam=True, alpha=alpha, beta=beta, mu=mu)
    f1, _ = o.umrsf(d3, alpha, 1.0)
    o.set_attenuation(mu)
    f2, _ = o.umrsf(d3, beta, 0.0, cur_pass=2)
    o.set_attenuation(0.0)
    assert np.abs(c.f3 - (f1 + f2)).max() < 1e-10 and np.abs(f2[:, 10]).max() > 1e-5 and np.abs(f2[:, :10]).max() == 0.0


def test_c_host_program(oracle_mod, tmp_path):
    """A C99 program (tests/c/routec_driver.c) drives the boundary: routec_fock_jk by reference on a registered context,
    oqpb_fock, and -- with two GPUs visible -- the multi-device context.  Its Fock matrices are compared with the oracle."""
    import os
    import struct
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "openqp_b200")
    exe = tmp_path / "routec_driver"
    r = subprocess.run(["gcc", "-std=c99", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "c", "routec_driver.c"),
                        "-o", str(exe), "-L", libdir, "-lopenqp_b200", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    bs = B.BasisSet(B.water_dimer(), "cc-pvdz")
    d = np.ascontiguousarray(pack(decaying_density(bs)))
    dump = tmp_path / "dump.bin"
    with open(dump, "wb") as f:
        f.write(struct.pack("4i", bs.nshell, bs.nprim, bs.nbf, 1 if bs.spherical else 0))
        for a in (bs.am, bs.harmonic, bs.ncontr, bs.g_offset, bs.ao_offset, bs.naos):
            f.write(np.ascontiguousarray(a, dtype=np.int32).tobytes())
        for a in (bs.ex, bs.cc, bs.centers, d):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    out = tmp_path / "out.bin"
    r = subprocess.run([str(exe), str(dump), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(out, "rb").read()
    nskipped, ndev = struct.unpack("qi", raw[:12])
    f3 = np.frombuffer(raw[12:], dtype=np.float64).reshape(3, bs.ntri)
    o = oracle_mod.Oracle(bs)
    o.set_screening()
    fo, st = o.fock(d)
    assert np.abs(f3[0] - fo[0]).max() < FOCK_TOL  # routec_fock_jk
    assert np.abs(f3[1] - fo[0]).max() < FOCK_TOL  # oqpb_fock
    assert np.abs(f3[2] - fo[0]).max() < FOCK_TOL  # multi-device context (or a copy on a 1-GPU box)
    assert abs(nskipped - st["nschwz"]) <= 2       # device Schwarz matrix: ulp-level differences only
    print(r.stdout.strip())
