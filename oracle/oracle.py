"""ctypes wrapper over oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product path (openqp_b200/*) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

RHF, UROHF, TD, MRSF, TDGRD, RPAGRD, UMRSF = 0, 1, 2, 3, 5, 6, 7


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle_int2.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.orc_create.restype = C.c_void_p
        L.orc_npairs_prim.restype = C.c_long
        L.orc_quartet_list.restype = C.c_long
        L.orc_eri_block.restype = C.c_int
        L.orc_max_threads.restype = C.c_int
    return _LIB


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


class Oracle:
    """Mirrors int2_compute_t (int2.F90:137-185): init -> set_screening -> run(consumer)."""

    def __init__(self, basis, cutoff: float = 5e-11):
        self.basis = basis
        L = lib()
        b = basis
        self.h = C.c_void_p(L.orc_create(
            C.c_int(b.nshell), _p(b.am, C.c_int), _p(b.ncontr, C.c_int), _p(b.g_offset, C.c_int),
            _p(b.ao_offset, C.c_int), _p(b.naos, C.c_int), _p(b.harmonic, C.c_int), _p(b.ex), _p(b.cc),
            _p(b.centers), C.c_int(1 if b.spherical else 0)))
        self.set_cutoff(cutoff)
        self.schwarz = None

    def __del__(self):
        try:
            lib().orc_destroy(self.h)
        except Exception:
            pass

    def set_cutoff(self, cutoff: float):
        self.cutoff = cutoff
        lib().orc_set_cutoff(self.h, C.c_double(cutoff))

    def set_screening(self, schwarz=None):
        ns = self.basis.nshell
        if schwarz is None:
            q = np.zeros((ns, ns))
            lib().orc_schwarz(self.h, _p(q))
        else:
            q = np.ascontiguousarray(schwarz, dtype=np.float64)
            lib().orc_set_schwarz(self.h, _p(q))
        self.schwarz = q
        return q

    def eri_block(self, i, j, k, l):
        b = self.basis
        mx = max((x + 1) * (x + 2) // 2 for x in b.am)
        out = np.zeros(mx ** 4)
        n = np.zeros(4, dtype=np.int32)
        lib().orc_eri_block(self.h, C.c_int(i), C.c_int(j), C.c_int(k), C.c_int(l), _p(out), _p(n, C.c_int))
        return out[: int(np.prod(n))].reshape(tuple(int(x) for x in n)).copy()

    def dense_eri(self):
        n = self.basis.nbf
        eri = np.zeros((n, n, n, n))
        lib().orc_dense_eri(self.h, _p(eri))
        return eri

    def _run(self, kind, d, nfocks, ncomp, se, sc, flags, f, f2, nthreads, pair_lo, pair_hi, stride, offset):
        st = np.zeros(3, dtype=np.int64)
        lib().orc_run(self.h, C.c_int(kind), _p(d), C.c_int(nfocks), C.c_int(ncomp), C.c_double(se), C.c_double(sc),
                      C.c_int(flags), _p(f), _p(f2) if f2 is not None else None, C.c_int(nthreads),
                      C.c_long(pair_lo), C.c_long(pair_hi), C.c_int(stride), C.c_int(offset), _p(st, C.c_long))
        return {"nschwz": int(st[0]), "nquartets": int(st[1]), "nints": int(st[2])}

    def fock(self, d_packed, scale_exchange=1.0, scale_coulomb=1.0, urohf=False, nthreads=0, post=True,
             pair_lo=0, pair_hi=-1, stride=1, offset=0):
        """fock_jk (scf_addons.F90:1063-1214): d,f packed (nfocks, ntri); returns (f, stats)."""
        d = np.ascontiguousarray(np.atleast_2d(d_packed), dtype=np.float64)
        nf = d.shape[0]
        f = np.zeros_like(d)
        st = self._run(UROHF if urohf else RHF, d, nf, 0, scale_exchange, scale_coulomb, 0, f, None, nthreads,
                       pair_lo, pair_hi, stride, offset)
        if post:
            lib().orc_fock_post(C.c_int(self.basis.nbf), C.c_int(nf), _p(f))
        return f, st

    def set_attenuation(self, mu: float):
        """switch to Erf-attenuated integrals erf(mu r)/r and the attenuated Schwarz matrix (mu > 0) or back (mu = 0)"""
        lib().orc_set_attenuation(self.h, C.c_double(mu))

    def active_schwarz(self):
        ns = self.basis.nshell
        q = np.zeros((ns, ns))
        lib().orc_get_schwarz(self.h, _p(q))
        return q

    def schwarz_attenuated(self, mu: float):
        """Schwarz matrix of the Erf-attenuated integrals (ints_exchange with mu2, int2.F90:674-681)"""
        self.set_attenuation(mu)
        q = self.active_schwarz()
        self.set_attenuation(0.0)
        return q

    def fock_cam(self, d_packed, alpha, beta, mu, alpha_coulomb=1.0, beta_coulomb=0.0, urohf=False, nthreads=0,
                 stride=1, offset=0):
        """int2_run_cam (int2.F90:538-584): pass 1 regular integrals with (scale_coulomb, scale_exchange) =
        (alpha_coulomb, alpha); pass 2 attenuated integrals with (beta_coulomb, beta); both into the same Fock, then the
        fock_jk post-scaling.  Returns (f, stats of pass 2) -- `skipped` is what the second run_generic leaves."""
        f1, st1 = self.fock(d_packed, alpha, alpha_coulomb, urohf=urohf, nthreads=nthreads, post=False, stride=stride, offset=offset)
        self.set_attenuation(mu)
        try:
            f2, st = self.fock(d_packed, beta, beta_coulomb, urohf=urohf, nthreads=nthreads, post=False, stride=stride, offset=offset)
        finally:
            self.set_attenuation(0.0)
        f = np.ascontiguousarray(f1 + f2)
        lib().orc_fock_post(C.c_int(self.basis.nbf), C.c_int(f.shape[0]), _p(f))
        st = dict(st, nquartets_both_passes=st1["nquartets"] + st["nquartets"])
        return f, st

    def td(self, d2, scale_exchange=1.0, scale_coulomb=1.0, int_apb=True, int_amb=False, tamm_dancoff=False,
           tamm_dancoff_coulomb=False, nthreads=0, post=True):
        """int2_td_data_t (tdhf_lib.F90:11-31): d2[v] = P_v as numpy (nvec, nbf, nbf) with P_v[mu,nu];
        returns apb, amb (nvec, nbf, nbf)."""
        nbf = self.basis.nbf
        d2 = np.asarray(d2, dtype=np.float64)
        nv = d2.shape[0]
        dF = np.ascontiguousarray(np.transpose(d2, (0, 2, 1)))  # Fortran (mu,nu,v): mu fastest
        apb = np.zeros_like(dF)
        amb = np.zeros_like(dF)
        flags = (1 if int_apb else 0) | (2 if int_amb else 0) | (4 if tamm_dancoff else 0) | (8 if tamm_dancoff_coulomb else 0)
        st = self._run(TD, dF, nv, 0, scale_exchange, scale_coulomb, flags, apb, amb, nthreads, 0, -1, 1, 0)
        if post:
            lib().orc_td_post(C.c_int(nbf), C.c_int(nv), _p(apb))
        return np.transpose(apb, (0, 2, 1)).copy(), np.transpose(amb, (0, 2, 1)).copy(), st

    def mrsf(self, d3, scale_exchange=1.0, scale_coulomb=1.0, nthreads=0, stride=1, offset=0, cur_pass=1):
        """int2_mrsf_data_t (tdhf_mrsf_lib.F90:8-26): d3 numpy (nvec, ncomp, nbf, nbf) [v,c,mu,nu];
        returns f3 same shape."""
        d3 = np.asarray(d3, dtype=np.float64)
        nv, nc, nbf, _ = d3.shape
        dF = np.ascontiguousarray(np.transpose(d3, (3, 2, 1, 0)))  # Fortran d3(v,c,mu,nu): v fastest
        f3 = np.zeros_like(dF)
        st = self._run(MRSF, dF, nv, nc, scale_exchange, scale_coulomb, 16 if cur_pass == 2 else 0, f3, None, nthreads, 0, -1,
                       stride, offset)
        return np.transpose(f3, (3, 2, 1, 0)).copy(), st

    def tdgrd(self, d2, scale_exchange=1.0, scale_coulomb=1.0, int_apb=True, int_amb=False, nthreads=0):
        """int2_tdgrd_data_t (tdhf_lib.F90:33-36, 228-295): d2 numpy (2, nbf, nbf) = two spin blocks P_s[mu, nu];
        returns apb (2, nbf, nbf) symmetrised as parallel_stop does, amb (2, nbf, nbf) (second block stays zero)."""
        nbf = self.basis.nbf
        d2 = np.asarray(d2, dtype=np.float64)
        assert d2.shape == (2, nbf, nbf)
        dF = np.ascontiguousarray(np.transpose(d2, (0, 2, 1)))
        apb = np.zeros_like(dF)
        amb = np.zeros_like(dF)
        flags = (1 if int_apb else 0) | (2 if int_amb else 0)
        st = self._run(TDGRD, dF, 2, 0, scale_exchange, scale_coulomb, flags, apb, amb, nthreads, 0, -1, 1, 0)
        lib().orc_td_post(C.c_int(nbf), C.c_int(2), _p(apb))
        return np.transpose(apb, (0, 2, 1)).copy(), np.transpose(amb, (0, 2, 1)).copy(), st

    def rpagrd(self, xpy=None, xmy=None, t=None, nspin=1, scale_exchange=1.0, scale_coulomb=1.0, nthreads=0):
        """int2_rpagrd_data_t (tdhf_lib.F90:42-57, 1068-1320): inputs numpy (n, nspin, nbf, nbf) [q, s, mu, nu] or None;
        returns hpp, hpt (symmetrised), hmm with the shapes of xpy, t, xmy."""
        nbf = self.basis.nbf

        def prep(a):
            if a is None:
                return np.zeros((0, nspin, nbf, nbf)), 0
            a = np.asarray(a, dtype=np.float64)
            assert a.shape[1:] == (nspin, nbf, nbf)
            return np.ascontiguousarray(np.transpose(a, (0, 1, 3, 2))), a.shape[0]  # Fortran (mu, nu, s, q)

        X, npp = prep(xpy)
        M, nm = prep(xmy)
        T, nt = prep(t)
        hpp, hpt, hmm = np.zeros_like(X), np.zeros_like(T), np.zeros_like(M)
        st = np.zeros(3, dtype=np.int64)
        lib().orc_run_rpagrd(self.h, C.c_int(nspin), C.c_int(npp), C.c_int(nm), C.c_int(nt), _p(X) if npp else None,
                             _p(M) if nm else None, _p(T) if nt else None, C.c_double(scale_exchange), C.c_double(scale_coulomb),
                             _p(hpp) if npp else None, _p(hpt) if nt else None, _p(hmm) if nm else None, C.c_int(nthreads),
                             _p(st, C.c_long))
        tr = lambda a: np.transpose(a, (0, 1, 3, 2)).copy()
        return tr(hpp), tr(hpt), tr(hmm), {"nschwz": int(st[0]), "nquartets": int(st[1]), "nints": int(st[2])}

    def umrsf(self, d3, scale_exchange=1.0, scale_coulomb=1.0, nthreads=0, cur_pass=1):
        """int2_umrsf_data_t (tdhf_mrsf_lib.F90:28-32, 337-426): d3 numpy (nvec, 11, nbf, nbf); returns f3 likewise."""
        d3 = np.asarray(d3, dtype=np.float64)
        nv, nc, nbf, _ = d3.shape
        assert nc == 11
        dF = np.ascontiguousarray(np.transpose(d3, (3, 2, 1, 0)))
        f3 = np.zeros_like(dF)
        st = self._run(UMRSF, dF, nv, nc, scale_exchange, scale_coulomb, 16 if cur_pass == 2 else 0, f3, None, nthreads, 0, -1, 1, 0)
        return np.transpose(f3, (3, 2, 1, 0)).copy(), st

    def mrsf_cam(self, d3, alpha, beta, mu, alpha_coulomb=1.0, beta_coulomb=0.0, nthreads=0):
        """int2_run_cam with the MRSF consumer: pass 1 regular (all components), pass 2 attenuated integrals, exchange of
        component 7 only (tdhf_mrsf_lib.F90:312-326)."""
        f1, _ = self.mrsf(d3, alpha, alpha_coulomb, nthreads=nthreads)
        self.set_attenuation(mu)
        try:
            f2, st = self.mrsf(d3, beta, beta_coulomb, nthreads=nthreads, cur_pass=2)
        finally:
            self.set_attenuation(0.0)
        return f1 + f2, st

    def td_cam(self, d2, alpha, beta, mu, alpha_coulomb=1.0, beta_coulomb=0.0, nthreads=0, **kw):
        """int2_run_cam with the TD consumer: the same update in both passes with the pass's scale factors
        (tdhf_lib.F90:140-224); the stop-time symmetrisation of apb is applied once to the sum."""
        a1, b1, _ = self.td(d2, alpha, alpha_coulomb, nthreads=nthreads, post=False, **kw)
        self.set_attenuation(mu)
        try:
            a2, b2, st = self.td(d2, beta, beta_coulomb, nthreads=nthreads, post=False, **kw)
        finally:
            self.set_attenuation(0.0)
        apb = np.ascontiguousarray(np.transpose(a1 + a2, (0, 2, 1)))
        lib().orc_td_post(C.c_int(self.basis.nbf), C.c_int(apb.shape[0]), _p(apb))
        return np.transpose(apb, (0, 2, 1)).copy(), b1 + b2, st

    def quartet_list(self, d_packed, want_list=True):
        d = np.ascontiguousarray(np.atleast_2d(d_packed), dtype=np.float64)
        ns = C.c_long(0)
        n = lib().orc_quartet_list(self.h, _p(d), C.c_int(d.shape[0]), None, C.c_long(0), C.byref(ns))
        out = None
        if want_list:
            out = np.zeros((n, 4), dtype=np.int32)
            lib().orc_quartet_list(self.h, _p(d), C.c_int(d.shape[0]), _p(out, C.c_int), C.c_long(n), C.byref(ns))
        return out, int(n), int(ns.value)

    def pair_order(self):
        """cost-sorted bra shell-pair list (int2.F90:864-921) as (npairs, 2) [i, j], i >= j, 0-based"""
        L = lib()
        L.orc_pair_order.restype = C.c_long
        n = L.orc_pair_order(self.h, None, None)
        pi = np.zeros(n, dtype=np.int32)
        pj = np.zeros(n, dtype=np.int32)
        L.orc_pair_order(self.h, _p(pi, C.c_int), _p(pj, C.c_int))
        return np.stack([pi, pj], axis=1)

    def sample_mask(self, stride, offset):
        """byte mask over canonical pair ids i(i+1)/2+j of the bra pairs a (stride, offset) run visits"""
        po = self.pair_order()
        sel = ((np.arange(len(po)) + 1) % stride) == offset if stride > 1 else np.ones(len(po), bool)
        mask = np.zeros(len(po), dtype=np.uint8)
        ij = po[sel].astype(np.int64)
        mask[ij[:, 0] * (ij[:, 0] + 1) // 2 + ij[:, 1]] = 1
        return mask

    def shlden(self, kind, d, nfocks, ncomp=0):
        ns = self.basis.nshell
        dsh = np.zeros((ns, ns))
        mx = C.c_double(0)
        d = np.ascontiguousarray(d, dtype=np.float64)
        lib().orc_shlden(self.h, C.c_int(kind), _p(d), C.c_int(nfocks), C.c_int(ncomp), _p(dsh), C.byref(mx))
        return dsh, mx.value

    def int1e(self):
        b = self.basis
        n = b.nbf
        S, T, V = np.zeros((n, n)), np.zeros((n, n)), np.zeros((n, n))
        Z = np.ascontiguousarray(b.mol.Z, dtype=np.float64)
        xyz = np.ascontiguousarray(b.mol.xyz, dtype=np.float64)
        lib().orc_int1e(self.h, C.c_int(b.mol.natom), _p(Z), _p(xyz), _p(S), _p(T), _p(V))
        return S, T, V


def rys(nroots: int, x: float):
    u = np.zeros(nroots)
    w = np.zeros(nroots)
    lib().orc_rys(C.c_int(nroots), C.c_double(x), _p(u), _p(w))
    return u, w


def set_fast_rys(on: bool):
    """Timed CPU baseline only: roots/weights of nroots <= 7 from Chebyshev tables instead of the general Stieltjes
    algorithm (the reference itself uses polynomial fits for nroots <= 5, rys.F90:45-2695).  Off for every parity test."""
    lib().orc_set_fast_rys(C.c_int(1 if on else 0))


def max_threads() -> int:
    return lib().orc_max_threads()
