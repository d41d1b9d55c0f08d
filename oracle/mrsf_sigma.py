"""CPU restatement (numpy) of the reference's MRSF sigma step -- TEST INFRASTRUCTURE, not shipped.

The reference's Davidson loop (source/modules/tdhf_mrsf_energy.F90:690-850) turns every new trial vector X (MO
occ-alpha x virt-beta parameterisation) into (A-B) X with the triple
    6a  iatogen + mrsfcbc      MO amplitudes -> seven AO "densities"          tdhf_lib.F90:480-498, tdhf_mrsf_lib.F90:940-1273
    6b  int2_mrsf_data_t run   J/K-type contractions of the seven densities  tdhf_mrsf_lib.F90:218-333 (the hot path)
    6c  mrsfmntoia + mrsfesum  AO Fock-like matrices -> MO, + orbital part   tdhf_mrsf_lib.F90:1463-1735, 1918-2036
and the device sigma session `routec_sig_init / _set_scale / _iter / _free` (source/modules/routec_sig.F90:28-56) replaces
the whole triple.  This module restates 6a and 6c line by line (0-based indices; lr1 = nocca-2, lr2 = nocca-1 are the two
singly occupied orbitals O1, O2) and strings them together around any J/K builder with int2_mrsf_data_t semantics, plus the
Guest-Saunders ROHF of scf.F90:1814-1893 needed to produce the reference state.

Pinned against the reference itself: tests/test_oracle_golden.py::test_mrsf_ch2o_golden reproduces the ROHF energy and the
three MRSF-CIS triplet excitation energies of examples/MRSF-TDDFT/CH2O_MRSFTDDFT_SYMMETRY_BLOCK_COVERAGE.json (no
functional: pure two-electron response), which pins int2_mrsf_data_t of the oracle to the real binary.
"""
from __future__ import annotations

import numpy as np

ISQRT2 = 1.0 / np.sqrt(2.0)


def iatogen(pv, nocca, noccb, nbf):
    """tdhf_lib.F90:480-498: pv(nocca, nbf-noccb) column-major (i fastest) -> av(nbf, nbf), zero elsewhere."""
    av = np.zeros((nbf, nbf))
    av[:nocca, noccb:] = np.asarray(pv).reshape((nbf - noccb, nocca)).T
    return av


def gentoia(wrk, nocca, noccb):
    """the packing loops at the end of mrsfmntoia / mrsfesum (ij runs over j = noccb+1..nbf outer, i = 1..nocca inner)."""
    return np.ascontiguousarray(wrk[:nocca, noccb:].T).ravel()


def mrsfcbc(va, vb, bvec, nocca, noccb, mrst):
    """tdhf_mrsf_lib.F90:940-1273.  Returns fmrsf[7, nbf, nbf] in the reference's component order
    (bo2v, bo1v, bco1, bco2, o21v, co12, ball), fmrsf[c, mu, nu]."""
    nbf = va.shape[0]
    lr1, lr2 = nocca - 2, nocca - 1
    f = np.zeros((7, nbf, nbf))
    bo2v, bo1v, bco1, bco2, o21v, co12, ball = f
    tv2 = vb[:, nocca:] @ bvec[lr2, nocca:]  # sum_a C^b(mu,a) X(O2,a)       :1001-1004, 1119-1122
    tv1 = vb[:, nocca:] @ bvec[lr1, nocca:]  # sum_a C^b(mu,a) X(O1,a)       :1028-1031, 1130-1133
    bo2v += np.outer(va[:, lr2], tv2)        # :1008-1011
    bo1v += np.outer(va[:, lr1], tv1)        # :1035-1038
    if noccb > 0:
        tc1 = va[:, :noccb] @ bvec[:noccb, lr1]  # sum_i C^a(mu,i) X(i,O1)   :1061-1064
        tc2 = va[:, :noccb] @ bvec[:noccb, lr2]  # sum_i C^a(mu,i) X(i,O2)   :1088-1091
        bco1 += np.outer(tc1, vb[:, lr1])        # :1068-1071
        bco2 += np.outer(tc2, vb[:, lr2])        # :1095-1098
    o21v += np.outer(tv2, va[:, lr1]) - np.outer(tv1, va[:, lr2])  # :1124-1138
    if noccb > 0:
        co12 += np.outer(vb[:, lr2], tc1) - np.outer(vb[:, lr1], tc2)  # :1141-1161
    ball += bo2v + bo1v + bco1 + bco2  # :1164
    if noccb > 0:
        tmp = vb[:, nocca:] @ bvec[:noccb, nocca:].T  # (nbf, noccb)    :1167-1170
        ball += va[:, :noccb] @ tmp.T                 # :1172-1175
    if mrst == 1:  # :1178-1186
        ball += (np.outer(va[:, lr2], vb[:, lr1]) * bvec[lr2, lr1] + np.outer(va[:, lr1], vb[:, lr2]) * bvec[lr1, lr2]
                 + (np.outer(va[:, lr1], vb[:, lr1]) - np.outer(va[:, lr2], vb[:, lr2])) * bvec[lr1, lr1] * ISQRT2)
    elif mrst == 3:  # :1187-1193
        ball += (np.outer(va[:, lr1], vb[:, lr1]) + np.outer(va[:, lr2], vb[:, lr2])) * bvec[lr1, lr1] * ISQRT2
    return f


def mrsfmntoia(fmrsf, va, vb, noca, nocb, mrst):
    """tdhf_mrsf_lib.F90:1463-1735: the seven AO matrices of one vector (fmrsf[c, mu, nu]) -> MO amplitudes."""
    ado2v, ado1v, adco1, adco2, ao21v, aco12, agdlr = fmrsf
    lr1, lr2 = noca - 2, noca - 1
    scr = va.T @ agdlr @ vb  # :1604-1614
    wrk = scr.copy()
    if noca > 2:
        tmp = ado1v @ vb[:, lr2] + aco12 @ vb[:, lr1]        # :1617-1624
        wrk[:noca - 2, lr2] += va[:, :noca - 2].T @ tmp      # :1626-1630
        tmp = ado2v @ vb[:, lr1] - aco12 @ vb[:, lr2]        # :1633-1640
        wrk[:noca - 2, lr1] += va[:, :noca - 2].T @ tmp      # :1642-1646
    tmp = adco2.T @ va[:, lr1] + ao21v.T @ va[:, lr2]        # :1649-1656
    wrk[lr1, noca:] += vb[:, noca:].T @ tmp                  # :1657-1660
    tmp = adco1.T @ va[:, lr2] - ao21v.T @ va[:, lr1]        # :1663-1670
    wrk[lr2, noca:] += vb[:, noca:].T @ tmp                  # :1671-1674
    if mrst == 1:  # :1677-1679
        wrk[lr1, lr1] = (scr[lr1, lr1] - scr[lr2, lr2]) * ISQRT2
        wrk[lr2, lr2] = 0.0
    elif mrst == 3:  # :1680-1685
        wrk[lr1, lr1] = (scr[lr1, lr1] + scr[lr2, lr2]) * ISQRT2
        wrk[lr2, lr1] = wrk[lr1, lr2] = wrk[lr2, lr2] = 0.0
    return gentoia(wrk, noca, nocb)


def mrsfesum(wrk, fij, fab, nocca, noccb, mrst):
    """tdhf_mrsf_lib.F90:1918-2036: the orbital-energy (one-electron) part of (A-B) X; wrk = iatogen(X),
    fij / fab = alpha / beta Fock matrices in the MO basis.  Returns the amplitudes to ADD to mrsfmntoia's."""
    nbf = wrk.shape[0]
    lr1, lr2 = nocca - 2, nocca - 1
    scr = wrk.copy()
    scr[lr1, lr1] = scr[lr2, lr2] = 0.0
    tmp1 = np.zeros((nbf, nbf))
    tmp1[:nocca, noccb:] = scr[:nocca, noccb:] @ fab[noccb:, noccb:].T - fij[:nocca, :nocca] @ scr[:nocca, noccb:]  # :1951-1960
    xlr = wrk[lr1, lr1]
    wrk1 = np.zeros((nbf, nbf))
    wrk1[:nocca, noccb:] = tmp1[:nocca, noccb:]
    s2 = 1.0 if mrst == 3 else -1.0  # sign of the O2 terms: singlet (-), triplet (+)   :1964-2011
    wrk1[lr1, noccb:] += fab[noccb:, lr1] * xlr * ISQRT2
    wrk1[lr2, noccb:] += s2 * fab[noccb:, lr2] * xlr * ISQRT2
    wrk1[:nocca, lr1] -= fij[:nocca, lr1] * xlr * ISQRT2
    wrk1[:nocca, lr2] -= s2 * fij[:nocca, lr2] * xlr * ISQRT2
    dumn = (-fij[lr1, :nocca] @ scr[:nocca, lr1] - s2 * (fij[lr2, :nocca] @ scr[:nocca, lr2])
            + fab[lr1, noccb:] @ scr[lr1, noccb:] + s2 * (fab[lr2, noccb:] @ scr[lr2, noccb:]))
    wrk1[lr1, lr1] = dumn * ISQRT2 + xlr * (fab[lr1, lr1] + fab[lr2, lr2] - fij[lr1, lr1] - fij[lr2, lr2]) * 0.5
    if mrst == 1:
        wrk1[lr2, lr2] = 0.0
    elif mrst == 3:
        wrk1[lr2, lr1] = wrk1[lr1, lr2] = wrk1[lr2, lr2] = 0.0
    return gentoia(wrk1, nocca, noccb)


def excluded_amplitudes(nocca, noccb, nbf, mrst):
    """amplitude indices the sigma step zeroes (their unit vectors span the null space the Davidson never visits:
    xvec_dim-1 / xvec_dim-3 in tdhf_mrsf_energy.F90:306-311)."""
    lr1, lr2 = nocca - 2, nocca - 1
    idx = lambda i, j: (j - noccb) * nocca + i
    return [idx(lr2, lr2)] if mrst == 1 else [idx(lr2, lr1), idx(lr1, lr2), idx(lr2, lr2)]


def sigma(bvecs, va, vb, fa, fb, nocca, noccb, mrst, jk_mrsf, scale_exchange=1.0):
    """(A-B) X for the columns of bvecs (ntrial, nv): tdhf_mrsf_energy.F90:690-826 (native path).
    jk_mrsf(d3[v, 7, mu, nu], scale_exchange) -> f3 of the same shape (int2_mrsf_data_t with tamm_dancoff = .true.)."""
    nbf = va.shape[0]
    bvecs = np.asarray(bvecs, dtype=np.float64)
    nv = bvecs.shape[1]
    X = [iatogen(bvecs[:, k], nocca, noccb, nbf) for k in range(nv)]
    d3 = np.stack([mrsfcbc(va, vb, X[k], nocca, noccb, mrst) for k in range(nv)])
    f3 = np.array(jk_mrsf(d3, scale_exchange))
    if mrst == 3:
        f3[:, :6] = -f3[:, :6]  # :762-763
    out = np.zeros_like(bvecs)
    for k in range(nv):
        out[:, k] = mrsfmntoia(f3[k], va, vb, nocca, noccb, mrst) + mrsfesum(X[k], fa, fb, nocca, noccb, mrst)
    return out


def rohf(nbf, S, H, enuc, fock2e_urohf, nalpha, nbeta, maxit=200, conv=1e-11, verbose=False):
    """Guest-Saunders ROHF (scf.F90:799-816, form_rohf_fock :1814-1893, all six coupling coefficients 1/2).
    fock2e_urohf(d_packed[2, ntri]) -> f_packed[2, ntri] has fock_jk / int2_urohf_data_t semantics.
    Returns (E, C, eps, Fa_ao, Fb_ao): C the ROHF orbitals (mo_a == mo_b), Fa / Fb the spin Fock matrices the MRSF driver
    transforms to the MO basis (rohf_bak, scf.F90:1288-1290)."""
    from openqp_b200.scf import pack, unpack

    s, U = np.linalg.eigh(S)
    Xo = U @ np.diag(s ** -0.5) @ U.T
    na, nb = nalpha, nbeta

    def diag(F):
        e, Cp = np.linalg.eigh(Xo.T @ F @ Xo)
        return e, Xo @ Cp

    eps, C = diag(H)
    errs, focks = [], []
    e_old = 0.0
    for it in range(maxit):
        Da, Db = C[:, :na] @ C[:, :na].T, C[:, :nb] @ C[:, :nb].T
        f2 = fock2e_urohf(np.stack([pack(Da), pack(Db)]))
        Fa, Fb = H + unpack(f2[0], nbf), H + unpack(f2[1], nbf)
        e = enuc + 0.5 * (np.sum(Da * (H + Fa)) + np.sum(Db * (H + Fb)))
        fa, fb = C.T @ Fa @ C, C.T @ Fb @ C
        F = 0.5 * (fa + fb)
        F[:nb, nb:na] = fb[:nb, nb:na]; F[nb:na, :nb] = fb[nb:na, :nb]
        F[nb:na, na:] = fa[nb:na, na:]; F[na:, nb:na] = fa[na:, nb:na]
        SC = S @ C
        Fao = SC @ F @ SC.T
        Dt = Da + Db
        err = (Fao @ Dt @ S - S @ Dt @ Fao).ravel()
        if verbose:
            print(f"rohf it {it:3d} E = {e:.12f} err = {np.abs(err).max():.2e}")
        if abs(e - e_old) < conv and np.abs(err).max() < 1e-8:
            return e, C, eps, Fa, Fb
        e_old = e
        errs.append(err); focks.append(Fao)
        errs, focks = errs[-8:], focks[-8:]
        n = len(errs)
        Fd = Fao
        if n > 1:
            B = -np.ones((n + 1, n + 1)); B[n, n] = 0
            for a in range(n):
                for b in range(n):
                    B[a, b] = errs[a] @ errs[b]
            rhs = np.zeros(n + 1); rhs[n] = -1
            try:
                c = np.linalg.solve(B, rhs)[:n]
                Fd = sum(c[a] * focks[a] for a in range(n))
            except np.linalg.LinAlgError:
                pass
        eps, C = diag(Fd)
    raise RuntimeError("ROHF not converged")


def dense_response_matrix(va, vb, fa, fb, nocca, noccb, mrst, jk_mrsf, batch=64):
    """(A-B) on the amplitude space the Davidson works in (the zeroed amplitudes removed), by applying sigma to unit
    vectors; returns the symmetrised matrix."""
    nbf = va.shape[0]
    n = nocca * (nbf - noccb)
    keep = np.array([k for k in range(n) if k not in set(excluded_amplitudes(nocca, noccb, nbf, mrst))])
    A = np.zeros((n, len(keep)))
    for k0 in range(0, len(keep), batch):
        ks = keep[k0:k0 + batch]
        E = np.zeros((n, len(ks)))
        E[ks, np.arange(len(ks))] = 1.0
        A[:, k0:k0 + len(ks)] = sigma(E, va, vb, fa, fb, nocca, noccb, mrst, jk_mrsf)
    A = A[keep]
    return 0.5 * (A + A.T), float(np.abs(A - A.T).max())
