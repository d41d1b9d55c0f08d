"""CPU: pin the oracle (oracle/oracle_int2.cpp) against the reference's own golden vectors."""
import numpy as np
import pytest

from common import decaying_density, golden, random_sym_density, rpa_energies
from openqp_b200 import basis as B
from openqp_b200.scf import pack, scf, unpack


def _rhf(oracle_mod, mol, name, nocc, urohf_nbeta=None):
    bs = B.BasisSet(mol, name)
    o = oracle_mod.Oracle(bs)
    o.set_screening()
    S, T, V = o.int1e()
    if urohf_nbeta is None:
        e, D, F = scf(bs.nbf, S, T + V, mol.nuclear_repulsion(), lambda dp: o.fock(dp)[0], nocc)
    else:
        e, D, F = scf(bs.nbf, S, T + V, mol.nuclear_repulsion(), lambda dp: o.fock(dp, urohf=True)[0], nocc, urohf_nbeta)
    return e


@pytest.mark.parametrize("key,basis,tol", [("h2o_rhf_631gd", "6-31g(d)", 1e-8), ("h2o_rhf_631g", "6-31g", 1e-8),
                                           ("h2o_rhf_sto3g_nmr", "sto-3g", 1e-8)])
def test_h2o_rhf_golden_energy(oracle_mod, key, basis, tol):
    ref = golden("reference_energies.json")[key]["energy"]
    e = _rhf(oracle_mod, B.water(), basis, 5)
    assert abs(e - ref) < tol, (e, ref)


def test_h2o_uhf_triplet_golden_energy(oracle_mod):
    """examples/HF/H2O_UHF-HF_ENERGY.json pins int2_urohf_data_t (int2.F90:1488-1578)."""
    ref = golden("reference_energies.json")["h2o_uhf_triplet_631gd"]["energy"]
    e = _rhf(oracle_mod, B.water(), "6-31g(d)", 6, 4)
    assert abs(e - ref) < 1e-8, (e, ref)


def test_water_dimer_golden_energy(oracle_mod):
    ref = golden("reference_energies.json")["h2o_dimer_rhf_631gd"]["energy"]
    e = _rhf(oracle_mod, B.water_dimer(), "6-31g(d)", 10)
    assert abs(e - ref) < 1e-8, (e, ref)


def test_tdhf_golden_excitations(oracle_mod):
    """examples/TDHF/H2O_TDHF_ENERGY.json (RPA-TDHF/6-31G*) pins int2_td_data_t (tdhf_lib.F90:140-224)."""
    g = golden("reference_energies.json")["h2o_tdhf_631gd"]
    mol = B.water()
    bs = B.BasisSet(mol, "6-31g(d)")
    o = oracle_mod.Oracle(bs)
    o.set_screening()
    S, T, V = o.int1e()

    def td(P):
        apb, amb, _ = o.td(P, int_apb=True, int_amb=True)
        return apb, amb

    e, w = rpa_energies(bs, S, T + V, mol.nuclear_repulsion(), 5, lambda dp: o.fock(dp)[0], td)
    assert abs(e - g["energy"]) < 1e-8
    assert np.allclose(w, g["td_energies"], atol=2e-7), (w, g["td_energies"])


def test_pure_tables_match_reference():
    """Formula-generated projection tables (tools/gen_pure_tables.py) vs the reference's generated tables
    (int2_pure_generated.F90:119-210), committed as tests/golden/reference_pure_tables.json."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location(
        "gen_pure", os.path.join(os.path.dirname(__file__), "..", "tools", "gen_pure_tables.py"))
    gp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gp)
    ref = golden("reference_pure_tables.json")
    for l in (2, 3, 4):
        mine = {(c + 1, o + 1): v for c, row in enumerate(gp.table(l)) for o, v in row}
        theirs = {(a, b): v for a, b, v in ref[str(l)]}
        assert set(mine) == set(theirs)
        for k in mine:
            assert abs(mine[k] - theirs[k]) < 5e-15


def test_oracle_consumers_vs_dense_eri(oracle_mod):
    """storeints + every consumer update against plain einsum contractions of the dense ERI tensor
    (dumped with the semantics of modules/int2e.F90:181-194); spherical d and f shells included."""
    mol = B.water()
    bs = B.BasisSet(mol, "cc-pvtz")
    o = oracle_mod.Oracle(bs, cutoff=1e-14)
    o.set_screening()
    eri = o.dense_eri()
    assert np.abs(eri - eri.transpose(1, 0, 2, 3)).max() == 0
    assert np.abs(eri - eri.transpose(2, 3, 0, 1)).max() == 0
    n = bs.nbf
    D = random_sym_density(n, 5)
    J = np.einsum("abcd,cd->ab", eri, D)
    K = np.einsum("acbd,cd->ab", eri, D)
    f, _ = o.fock(pack(D))
    assert np.abs(unpack(f[0], n) - (J - 0.5 * K)).max() < 1e-11
    rng = np.random.default_rng(3)
    Da, Db = random_sym_density(n, 8), random_sym_density(n, 9)
    f, _ = o.fock(np.stack([pack(Da), pack(Db)]), urohf=True, scale_exchange=0.3, scale_coulomb=0.9)
    Jt = np.einsum("abcd,cd->ab", eri, Da + Db)
    for k, Ds in enumerate((Da, Db)):
        Ks = np.einsum("acbd,cd->ab", eri, Ds)
        assert np.abs(unpack(f[k], n) - (0.9 * Jt - 0.3 * Ks)).max() < 1e-11
    P = rng.normal(size=(2, n, n))
    apb, amb, _ = o.td(P, int_apb=True, int_amb=True)
    for v in range(2):
        Ps = P[v] + P[v].T
        assert np.abs(apb[v] - (2 * np.einsum("abcd,cd->ab", eri, Ps) - np.einsum("acbd,cd->ab", eri, Ps))).max() < 1e-10
        assert np.abs(amb[v] - np.einsum("acbd,cd->ab", eri, P[v].T - P[v])).max() < 1e-10
    d3 = rng.normal(size=(2, 7, n, n))
    f3, _ = o.mrsf(d3, 0.5, 0.7)
    for v in range(2):
        for c in range(7):
            ref = -0.5 * np.einsum("acbd,cd->ab", eri, d3[v, c])
            if c < 4:
                ref = ref + 0.7 * np.einsum("abcd,cd->ab", eri, d3[v, c])
            assert np.abs(f3[v, c] - ref).max() < 1e-10


def test_oracle_rys_against_mpmath(oracle_mod):
    """Rys roots/weights of the oracle (rys.F90:2697-2881 restated) against an independent 100-digit
    Gauss quadrature from Boys moments (tools/gen_rys_tables.py)."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location(
        "gen_rys", os.path.join(os.path.dirname(__file__), "..", "tools", "gen_rys_tables.py"))
    gr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gr)
    for R, xs in ((1, [0.0, 0.7, 12.0, 38.5]), (3, [0.2, 9.0, 52.0]), (5, [3.3, 64.0]), (7, [0.0, 21.0, 74.5, 90.0])):
        for x in xs:
            r, w = gr.rys(R, x)
            u, ww = oracle_mod.rys(R, x)
            order = np.argsort(u)  # the implicit-QL eigen-solver (rys.F90:2791-2881) does not sort its roots
            u, ww = u[order], ww[order]
            t2 = u / (1 + u)
            assert max(abs(t2[i] / float(r[i]) - 1) for i in range(R)) < 2e-13
            assert max(abs(ww[i] / float(w[i]) - 1) for i in range(R)) < 2e-13


def test_screening_counts_consistent(oracle_mod):
    """nschwz + survivors = all canonical quartets (int2.F90:756-805 bookkeeping)."""
    from common import decaying_density
    mol = B.benzene()
    bs = B.BasisSet(mol, "6-31g")
    o = oracle_mod.Oracle(bs)
    o.set_screening()
    d = pack(decaying_density(bs) * 1e-3)
    lst, n, nschwz = o.quartet_list(d)
    npair = bs.nshell * (bs.nshell + 1) // 2
    assert n + nschwz == npair * (npair + 1) // 2
    assert nschwz > 0 and n > 0
    f, st = o.fock(d)
    assert st["nschwz"] == nschwz and st["nquartets"] == n


def test_oracle_attenuated_integrals_closed_form(oracle_mod):
    """Erf-attenuated integrals of the CAM second pass (int_rys.F90:179-181, 225-227): contracted (ss|ss) integrals
    against the closed form  sum_prims c_a c_b c_c c_d 2 pi^{5/2} / (zeta eta sqrt(ab')) K_ab K_cd F0(rho' |PQ|^2),
    ab' = zeta + eta + zeta eta / mu^2, rho' = zeta eta / ab'; and the limits mu -> infinity (regular integrals)."""
    from scipy.special import erf
    bs = B.BasisSet(B.water_dimer(), "6-31g")
    mu = 0.33
    o = oracle_mod.Oracle(bs)
    o.set_screening()
    sshells = [i for i in range(bs.nshell) if bs.am[i] == 0]

    def f0(x):
        return 1.0 if x < 1e-14 else 0.5 * np.sqrt(np.pi / x) * erf(np.sqrt(x))

    def prims(i):
        g = bs.g_offset[i] - (1 if bs.g_offset.min() == 1 else 0)
        return [(bs.ex[g + k], bs.cc[g + k]) for k in range(bs.ncontr[i])], np.asarray(bs.centers[i], dtype=float)

    def closed(i, j, k, l, mu):
        (pa, A), (pb, Bc), (pc, Cc), (pd, Dc) = prims(i), prims(j), prims(k), prims(l)
        tot = 0.0
        for a, ca in pa:
            for b, cb in pb:
                z = a + b
                P = (a * A + b * Bc) / z
                kab = np.exp(-a * b / z * np.dot(A - Bc, A - Bc))
                for c, cc_ in pc:
                    for d, cd in pd:
                        e = c + d
                        Q = (c * Cc + d * Dc) / e
                        kcd = np.exp(-c * d / e * np.dot(Cc - Dc, Cc - Dc))
                        ab = z + e + (z * e / mu ** 2 if mu > 0 else 0.0)
                        rho = z * e / ab
                        tot += ca * cb * cc_ * cd * 2 * np.pi ** 2.5 / (z * e * np.sqrt(ab)) * kab * kcd * f0(rho * np.dot(P - Q, P - Q))
        return tot

    quartets = [(sshells[0], sshells[1], sshells[2], sshells[3]), (sshells[4], sshells[0], sshells[5], sshells[2]),
                (sshells[1], sshells[1], sshells[6], sshells[3])]
    regs = []
    for q in quartets:
        reg = o.eri_block(*q).ravel()[0]
        assert abs(reg - closed(*q, 0.0)) < 1e-12 * max(1.0, abs(reg))
        regs.append(reg)
    o.set_attenuation(mu)
    for q, reg in zip(quartets, regs):
        att = o.eri_block(*q).ravel()[0]
        assert abs(att - closed(*q, mu)) < 1e-12 * max(1.0, abs(att))
        assert 0 < att < reg  # erf(mu r)/r < 1/r
    o.set_attenuation(1.0e6)  # erf(mu r) -> 1
    for q in quartets:
        assert abs(o.eri_block(*q).ravel()[0] - closed(*q, 0.0)) < 1e-9
    o.set_attenuation(0.0)
    # CAM with beta = 0 is the regular build; the attenuated Schwarz matrix is bounded by the regular one
    d = pack(decaying_density(bs))
    f0_, _ = o.fock(d, 0.25, 1.0)
    f1_, _ = o.fock_cam(d, 0.25, 0.0, mu)
    assert np.abs(f0_ - f1_).max() < 1e-13
    qa = o.schwarz_attenuated(mu)
    assert (qa <= o.schwarz * (1 + 1e-12) + 1e-300).all() and qa.max() > 0


def test_oracle_strided_partitions_sum_to_full(oracle_mod):
    """The replicated-data split `mod(ij_pair, size) == rank` (int2.F90:759-761), which the bench's bounded CPU samples
    and the multi-GPU partition rely on: strided partial builds of every consumer add up to the full build."""
    bs = B.BasisSet(B.water_dimer(), "6-31g")
    o = oracle_mod.Oracle(bs)
    o.set_screening()
    d = pack(decaying_density(bs))
    full, st = o.fock(d, 0.5, 1.0, post=False)
    parts = [o.fock(d, 0.5, 1.0, post=False, stride=3, offset=r) for r in range(3)]
    assert np.abs(sum(p[0] for p in parts) - full).max() < 1e-12
    assert sum(p[1]["nquartets"] for p in parts) == st["nquartets"] and sum(p[1]["nschwz"] for p in parts) == st["nschwz"]
    rng = np.random.default_rng(2)
    d3 = rng.normal(size=(2, 7, bs.nbf, bs.nbf)) * 0.1
    f3, st3 = o.mrsf(d3, 0.5, 1.0)
    p3 = [o.mrsf(d3, 0.5, 1.0, stride=2, offset=r) for r in range(2)]
    assert np.abs(p3[0][0] + p3[1][0] - f3).max() < 1e-12
    assert p3[0][1]["nquartets"] + p3[1][1]["nquartets"] == st3["nquartets"]


def test_oracle_gradient_response_consumers_vs_dense_eri(oracle_mod):
    """The restatements of int2_tdgrd_data_t / int2_rpagrd_data_t / int2_umrsf_data_t (tdhf_lib.F90:228-295, 1068-1320;
    tdhf_mrsf_lib.F90:337-426) against plain J/K algebra on the dense ERI tensor."""
    from openqp_b200 import basis as B
    bs = B.BasisSet(B.water(), "6-31g(d)")
    o = oracle_mod.Oracle(bs, 1e-14)
    o.set_screening()
    eri = o.dense_eri()
    n = bs.nbf
    J = lambda P: np.einsum("abcd,cd->ab", eri, P)
    K = lambda P: np.einsum("abcd,bd->ac", eri, P)
    rng = np.random.default_rng(0)
    se, sc = 0.7, 0.9
    d2 = rng.normal(size=(2, n, n)) * 0.1
    apb, amb, _ = o.tdgrd(d2, se, sc, True, True)
    for s_ in range(2):
        assert np.abs(apb[s_] - (2 * sc * J(d2[0] + d2[1]) - se * K(d2[s_] + d2[s_].T))).max() < 1e-13
    assert np.abs(amb[0] - se * K(d2[0].T - d2[0])).max() < 1e-13 and np.abs(amb[1]).max() == 0.0
    sym = lambda a: a + np.swapaxes(a, -1, -2)
    xpy, t, xmy = sym(rng.normal(size=(2, 1, n, n)) * 0.1), sym(rng.normal(size=(1, 1, n, n)) * 0.1), rng.normal(size=(2, 1, n, n)) * 0.1
    hpp, hpt, hmm, _ = o.rpagrd(xpy, xmy, t, 1, se, sc)
    for q in range(2):
        assert np.abs(hpp[q, 0] - (4 * sc * J(xpy[q, 0]) - 2 * se * K(xpy[q, 0]))).max() < 1e-13
        assert np.abs(hmm[q, 0] - se * K(xmy[q, 0].T - xmy[q, 0])).max() < 1e-13
    assert np.abs(hpt[0, 0] - (4 * sc * J(t[0, 0]) - 2 * se * K(t[0, 0]))).max() < 1e-13
    xpy2, xmy2 = rng.normal(size=(1, 2, n, n)) * 0.1, rng.normal(size=(1, 2, n, n)) * 0.1
    hpp, _, hmm, _ = o.rpagrd(xpy2, xmy2, None, 2, se, sc)
    for s_ in range(2):
        assert np.abs(hpp[0, s_] - (2 * sc * J(xpy2[0, 0] + xpy2[0, 1]) - se * K(xpy2[0, s_] + xpy2[0, s_].T))).max() < 1e-13
    assert np.abs(hmm[0, 0] - se * K(xmy2[0, 0].T - xmy2[0, 0])).max() < 1e-13 and np.abs(hmm[0, 1]).max() == 0.0
    d3 = rng.normal(size=(2, 11, n, n)) * 0.1
    for cur_pass in (1, 2):
        f3, _ = o.umrsf(d3, se, sc, cur_pass=cur_pass)
        for v in range(2):
            for c in range(11):
                ref = (sc * J(d3[v, c]) if c < 8 else 0.0) - se * K(d3[v, c].T if c in (8, 9) else d3[v, c])
                if cur_pass == 2 and c < 10:
                    ref = 0.0 * ref
                assert np.abs(f3[v, c] - ref).max() < 1e-13


def _ch2o_rohf(oracle_mod):
    g = golden("reference_energies.json")["ch2o_mrsf_triplet_631g"]
    mol = B.Molecule(np.array(g["atoms"]), np.array(g["coord"]).reshape(-1, 3))
    bs = B.BasisSet(mol, "6-31g")
    o = oracle_mod.Oracle(bs, cutoff=1e-12)
    o.set_screening()
    S, T, V = o.int1e()
    nel = int(sum(g["atoms"]))
    na, nb = nel // 2 + 1, nel // 2 - 1
    from oracle import mrsf_sigma as MS
    e, C, eps, Fa, Fb = MS.rohf(bs.nbf, S, T + V, mol.nuclear_repulsion(), lambda dp: o.fock(dp, urohf=True)[0], na, nb)
    return g, bs, o, MS, e, C, C.T @ Fa @ C, C.T @ Fb @ C, na, nb


def test_mrsf_ch2o_golden(oracle_mod):
    """examples/MRSF-TDDFT/CH2O_MRSFTDDFT_SYMMETRY_BLOCK_COVERAGE.json: ROHF triplet + MRSF-CIS/6-31G triplet roots with NO
    functional -- the XC-free MRSF golden.  Pins int2_mrsf_data_t (tdhf_mrsf_lib.F90:218-333) and the sigma triple
    mrsfcbc / mrsfmntoia / mrsfesum (oracle/mrsf_sigma.py) to the real binary: the full (A-B) matrix is built by applying
    the sigma step to unit vectors and diagonalised.  The deck's JSON holds Davidson-converged roots (residual 1e-8 in
    ||r||^2, i.e. ~1e-6 Eh on a root); its header quotes the tighter nstate=20 run, which the dense solve reproduces."""
    g, bs, o, MS, e, C, fa, fb, na, nb = _ch2o_rohf(oracle_mod)
    assert abs(e - g["energy"]) < 1e-8, (e, g["energy"])
    A, asym = MS.dense_response_matrix(C, C, fa, fb, na, nb, 3, lambda d3, sx: o.mrsf(d3, scale_exchange=sx, scale_coulomb=sx)[0])
    assert asym < 1e-10  # (A-B) is symmetric on the Davidson's amplitude space
    w = np.linalg.eigvalsh(A)
    assert np.allclose(w[:4], g["roots_nstate20"], atol=2e-8), (w[:4], g["roots_nstate20"])
    assert np.allclose(w[:3], g["td_energies"], atol=5e-7), (w[:3], g["td_energies"])


MD_QUARTETS = [(9, 15, 31, 21), (9, 8, 31, 30), (9, 9, 15, 13), (15, 21, 7, 13), (31, 26, 21, 16), (7, 0, 26, 22), (8, 4, 30, 43)]


def test_oracle_eri_vs_mpmath_mcmurchie_davidson(oracle_mod):
    """SURVEY 8c cross-check: d / f shell quartets of (H2O)2 / cc-pVTZ (Cartesian 6d/10f, up to four distinct centres,
    contracted s / p partners) from an independent 40-digit McMurchie-Davidson evaluation (tests/md_eri.py: Hermite
    expansion + Boys function, no Rys quadrature anywhere) against shellquartet of the oracle.  Pins the f-shell and d-shell
    integral values, their normalisation (normalize_ints, int2.F90:1187-1207) and the reference's component order without
    the real binary; the Cartesian -> pure tables are pinned separately against the reference's own (test above)."""
    import md_eri
    bs = B.BasisSet(B.water_dimer(), "cc-pvtz", spherical=False)
    o = oracle_mod.Oracle(bs)
    o.set_screening()
    worst = 0.0
    for q in MD_QUARTETS:
        ref = np.array(md_eri.shell_quartet(bs, *q))
        i, j, k, l = q
        blk = o.eri_block(max(i, j), min(i, j), max(k, l), min(k, l))  # shellquartet takes canonical shell order
        if i < j:
            blk = blk.transpose(1, 0, 2, 3)
        if k < l:
            blk = blk.transpose(0, 1, 3, 2)
        assert blk.shape == ref.shape
        scale = np.abs(ref).max()
        assert scale > 1e-6, (q, scale)  # far above the primitive-pair cutoffs (int2_pairs.F90:245-249), which drop ~1e-10 terms
        err = np.abs(blk - ref).max() / scale
        worst = max(worst, err)
        assert err < 1e-12, (q, [int(bs.am[s]) for s in q], err)
    print(f"oracle vs McMurchie-Davidson/mpmath: {len(MD_QUARTETS)} quartets, worst relative block error {worst:.1e}")


def test_oracle_fast_rys_matches_general(oracle_mod):
    """The table-driven roots that only the TIMED CPU baseline switches on (orc_set_fast_rys; bench.py cpu_baseline and
    --impl reference) against the restated general algorithm: roots / weights to 1e-12 relative, a Fock build to 1e-12."""
    rng = np.random.default_rng(0)
    try:
        for R in range(1, 8):
            for x in np.concatenate([rng.uniform(0, 90, 60), [0.0, 1e-9, 38.999, 39.0, 74.999, 75.0, 200.0]]):
                oracle_mod.set_fast_rys(False)
                u0, w0 = oracle_mod.rys(R, float(x))
                oracle_mod.set_fast_rys(True)
                u1, w1 = oracle_mod.rys(R, float(x))
                o0, o1 = np.argsort(u0), np.argsort(u1)
                assert np.abs(u1[o1] / u0[o0] - 1).max() < 1e-12 and np.abs(w1[o1] / w0[o0] - 1).max() < 1e-12, (R, x)
        bs = B.BasisSet(B.water(), "cc-pvtz")
        o = oracle_mod.Oracle(bs)
        o.set_screening()
        d = pack(random_sym_density(bs.nbf, 3))
        oracle_mod.set_fast_rys(False)
        f0 = o.fock(d)[0]
        oracle_mod.set_fast_rys(True)
        f1 = o.fock(d)[0]
        assert np.abs(f1 - f0).max() < 1e-12 * max(1.0, np.abs(f0).max())
    finally:
        oracle_mod.set_fast_rys(False)


def test_oracle_attenuated_eri_vs_mpmath_mcmurchie_davidson(oracle_mod):
    """The Erf-attenuated integrals of the CAM second pass (int_rys.F90:179-181, 225-227: ab = zeta + eta + zeta eta / mu^2)
    for d / f quartets on up to four centres against the independent McMurchie-Davidson evaluation with the operator
    erf(mu r12) / r12 (tests/md_eri.py): pins the range-separated integrals beyond the contracted (ss|ss) closed form."""
    import md_eri
    bs = B.BasisSet(B.water_dimer(), "cc-pvtz", spherical=False)
    o = oracle_mod.Oracle(bs)
    o.set_screening()
    mu = 0.33  # CAM-B3LYP
    o.set_attenuation(mu)
    try:
        worst = 0.0
        for q in [(9, 15, 31, 21), (9, 9, 15, 13), (31, 26, 21, 16), (7, 0, 26, 22)]:
            i, j, k, l = q
            ref = np.array(md_eri.shell_quartet(bs, *q, mu=mu))
            blk = o.eri_block(max(i, j), min(i, j), max(k, l), min(k, l))
            if i < j:
                blk = blk.transpose(1, 0, 2, 3)
            if k < l:
                blk = blk.transpose(0, 1, 3, 2)
            err = np.abs(blk - ref).max() / np.abs(ref).max()
            worst = max(worst, err)
            assert err < 1e-12, (q, err)
        print(f"attenuated integrals, oracle vs McMurchie-Davidson/mpmath: worst relative block error {worst:.1e}")
    finally:
        o.set_attenuation(0.0)
