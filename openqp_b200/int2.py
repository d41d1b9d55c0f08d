"""Host-side mirror of the reference's two-electron driver interface, on top of the C ABI
(include/oqp_b200.h, libopenqp_b200.so).  The reference host is Fortran (no Fortran compiler in this image),
so this mirror is Python over ctypes; `fortran/oqp_b200_shim.F90` is the ISO_C_BINDING shim a maintainer
would compile into OpenQP.

Names follow /root/reference/source/integrals/int2.F90:137-185:
    Int2Compute.init / set_screening / set_cutoff / run(consumer) / clean, field `skipped`
    consumers Int2RhfData, Int2UrohfData (int2.F90:83-135), Int2TdData (tdhf_lib.F90:11-31),
    Int2MrsfData (tdhf_mrsf_lib.F90:8-26); `fock_jk` (scf_addons.F90:1063-1214).

There is NO CPU fallback: a missing library or a missing CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.environ.get("OQPB_LIB") or os.path.join(_HERE, "libopenqp_b200.so")  # OQPB_LIB: tuning variants (tools/)
_LIB = None

OQPB_TD_APB, OQPB_TD_AMB, OQPB_TD_TDA, OQPB_TD_TDA_COULOMB = 1, 2, 4, 8

_ERRORS = {1: "no CUDA device", 2: "bad argument", 3: "unsupported", 4: "call order", 5: "CUDA error"}


class Int2Error(RuntimeError):
    pass


def lib():
    """Load libopenqp_b200.so (built in-tree by openqp_b200.build); raises if it is missing."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(_LIBPATH):
            raise Int2Error(f"{_LIBPATH} not built: run `python -m openqp_b200.build` (no CPU fallback exists)")
        L = C.CDLL(_LIBPATH)
        L.oqpb_last_error.restype = C.c_char_p
        L.oqpb_last_flops.restype = C.c_double
        L.oqpb_last_kernel_ms.restype = C.c_double
        L.oqpb_fp64_peak_tflops.restype = C.c_double
        L.oqpb_get_quartets.restype = C.c_longlong
        L.oqpb_stream.restype = C.c_void_p
        for name in ("oqpb_last_error", "oqpb_last_flops", "oqpb_last_kernel_ms", "oqpb_fp64_peak_tflops",
                     "oqpb_ctx_destroy", "oqpb_stream", "oqpb_synchronize"):
            getattr(L, name).argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class Int2Compute:
    """int2_compute_t (int2.F90:137-185)."""

    def __init__(self, device: int = 0, ndevices: int = 1):
        """ndevices > 1: one context over GPUs 0 .. ndevices-1 of this node (oqpb_ctx_create_multi): the host-pointer
        entries split the work over the devices and sum the partial results with one NCCL all-reduce"""
        self._h = C.c_void_p()
        if ndevices > 1:
            rc = lib().oqpb_ctx_create_multi(C.byref(self._h), C.c_int(ndevices), None)
        else:
            rc = lib().oqpb_ctx_create(C.byref(self._h), C.c_int(device))
        if rc != 0:
            self._h = C.c_void_p()
            raise Int2Error(f"oqpb_ctx_create failed: {_ERRORS.get(rc, rc)} (this library has no CPU fallback)")
        self.device = device
        self.basis = None
        self.skipped = 0
        self.cutoff = None

    # -- error handling ---------------------------------------------------------------------------
    def _check(self, rc, what):
        if rc != 0:
            msg = lib().oqpb_last_error(self._h)
            raise Int2Error(f"{what}: {_ERRORS.get(rc, rc)}: {msg.decode() if msg else ''}")

    # -- int2_compute_t%init (int2.F90:245-289) -----------------------------------------------------
    def init(self, basis, cutoff: float = 5e-11):
        b = basis
        self.basis = b
        self._check(lib().oqpb_set_basis(
            self._h, C.c_int(b.nshell), C.c_int(b.nprim), _ip(b.am), _ip(b.harmonic), _ip(b.ncontr), _ip(b.g_offset),
            _ip(b.ao_offset), _ip(b.naos), _dp(b.ex), _dp(b.cc), _dp(b.centers), C.c_int(1 if b.spherical else 0)),
            "oqpb_set_basis")
        self.set_cutoff(cutoff)
        return self

    def set_cutoff(self, cutoff: float):
        self.cutoff = float(cutoff)
        self._check(lib().oqpb_set_cutoff(self._h, C.c_double(cutoff)), "oqpb_set_cutoff")

    # -- int2_compute_t%set_screening (int2.F90:473-477) -------------------------------------------
    def set_screening(self, schwarz=None):
        if schwarz is None:
            self._check(lib().oqpb_set_screening(self._h, None), "oqpb_set_screening")
        else:
            q = np.ascontiguousarray(schwarz, dtype=np.float64)
            assert q.shape == (self.basis.nshell, self.basis.nshell)
            self._check(lib().oqpb_set_screening(self._h, _dp(q)), "oqpb_set_screening")
        return self.schwarz()

    def schwarz(self):
        ns = self.basis.nshell
        q = np.zeros((ns, ns))
        self._check(lib().oqpb_get_schwarz(self._h, _dp(q)), "oqpb_get_schwarz")
        return q

    def set_partition(self, rank: int, nranks: int):
        self._check(lib().oqpb_set_partition(self._h, C.c_int(rank), C.c_int(nranks)), "oqpb_set_partition")

    def set_bra_mask(self, mask=None):
        """test hook (oqpb_set_bra_mask): restrict builds to quartets whose reference bra pair is flagged"""
        if mask is None:
            self._check(lib().oqpb_set_bra_mask(self._h, None, C.c_longlong(0)), "oqpb_set_bra_mask")
        else:
            m = np.ascontiguousarray(mask, dtype=np.uint8)
            self._check(lib().oqpb_set_bra_mask(self._h, m.ctypes.data_as(C.c_void_p), C.c_longlong(m.size)), "oqpb_set_bra_mask")

    def set_screening_cam(self, mu: float, schwarz_att=None):
        """Schwarz matrix of the Erf-attenuated integrals for the CAM second pass (int2.F90:674-685)"""
        if schwarz_att is None:
            self._check(lib().oqpb_set_screening_cam(self._h, C.c_double(mu), None), "oqpb_set_screening_cam")
        else:
            q = np.ascontiguousarray(schwarz_att, dtype=np.float64)
            self._check(lib().oqpb_set_screening_cam(self._h, C.c_double(mu), _dp(q)), "oqpb_set_screening_cam")
        ns = self.basis.nshell
        q = np.zeros((ns, ns))
        self._check(lib().oqpb_get_schwarz_cam(self._h, _dp(q)), "oqpb_get_schwarz_cam")
        return q

    # -- int2_compute_t%run (int2.F90:500-536, 589) -----------------------------------------------
    def run(self, consumer, cam=False, alpha=1.0, beta=0.0, mu=0.0, alpha_coulomb=1.0, beta_coulomb=0.0):
        """run(consumer): one pass with the consumer's scale factors.  run(consumer, cam=True, alpha, beta, mu): the
        range-separated two-pass build int2_run_cam (int2.F90:538-584) for every consumer."""
        if cam:
            if not hasattr(consumer, "_run_cam"):
                raise Int2Error("consumer without a CAM path")
            consumer._run_cam(self, alpha, beta, mu, alpha_coulomb, beta_coulomb)
        else:
            consumer._run(self)
        self.skipped = consumer.skipped
        return consumer

    def clean(self):
        if self._h:
            lib().oqpb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.clean()
        except Exception:
            pass

    # -- device-pointer entry points (bench / NCCL) -------------------------------------------------
    def fock_dev(self, d_ptr: int, f_ptr: int, nfocks: int, urohf=False, scale_exchange=1.0, scale_coulomb=1.0):
        self._check(lib().oqpb_fock_dev(self._h, C.c_int(1 if urohf else 0), C.c_void_p(d_ptr), C.c_void_p(f_ptr),
                                        C.c_int(nfocks), C.c_double(scale_exchange), C.c_double(scale_coulomb)),
                    "oqpb_fock_dev")

    def fock_cam_dev(self, d_ptr: int, f_ptr: int, nfocks: int, alpha, beta, mu, urohf=False, alpha_coulomb=1.0, beta_coulomb=0.0):
        """int2_run_cam (int2.F90:538-584) with device-resident packed d / f (raw accumulator; fock_post_dev scales)"""
        self._check(lib().oqpb_fock_cam_dev(self._h, C.c_int(1 if urohf else 0), C.c_void_p(d_ptr), C.c_void_p(f_ptr),
                                            C.c_int(nfocks), C.c_double(alpha), C.c_double(beta), C.c_double(mu),
                                            C.c_double(alpha_coulomb), C.c_double(beta_coulomb)), "oqpb_fock_cam_dev")

    def mrsf_dev(self, d3_ptr: int, f3_ptr: int, nvec: int, ncomp: int = 7, scale_exchange=1.0, scale_coulomb=1.0):
        """int2_mrsf_data_t with device-resident d3 / f3 (layout d3(v, c, mu, nu), v fastest)"""
        self._check(lib().oqpb_jk_mrsf_dev(self._h, C.c_void_p(d3_ptr), C.c_int(nvec), C.c_int(ncomp),
                                           C.c_double(scale_exchange), C.c_double(scale_coulomb), C.c_void_p(f3_ptr)),
                    "oqpb_jk_mrsf_dev")

    def fock_post_dev(self, f_ptr: int, nfocks: int):
        self._check(lib().oqpb_fock_post_dev(self._h, C.c_void_p(f_ptr), C.c_int(nfocks)), "oqpb_fock_post_dev")

    def synchronize(self):
        self._check(lib().oqpb_synchronize(self._h), "oqpb_synchronize")

    def set_stream(self, cuda_stream: int):
        self._check(lib().oqpb_set_stream(self._h, C.c_void_p(cuda_stream)), "oqpb_set_stream")

    def stream(self) -> int:
        return lib().oqpb_stream(self._h)

    # -- introspection --------------------------------------------------------------------------
    def last_stats(self):
        s = (C.c_longlong * 4)()
        lib().oqpb_last_stats(self._h, s)
        return {"nquartets": int(s[0]), "nschwz": int(s[1]), "launches": int(s[3]),
                "flops": lib().oqpb_last_flops(self._h), "kernel_ms": lib().oqpb_last_kernel_ms(self._h)}

    def profile(self, enable=True):
        """Enable/disable per-class timing; returns the table accumulated since the previous call."""
        out = np.zeros((55, 4))
        lib().oqpb_profile(self._h, C.c_int(1 if enable else 0), _dp(out))
        names = ["ss", "ps", "pp", "ds", "dp", "dd", "fs", "fp", "fd", "ff"]
        tab = {}
        for a in range(10):
            for b in range(a + 1):
                r = out[a * (a + 1) // 2 + b]
                if r[1] > 0:
                    tab[f"({names[a]}|{names[b]})"] = {"ms": r[0], "quartets": int(r[1]), "prims": int(r[2]), "flops": r[3]}
        return tab

    def record_quartets(self, enable=True):
        lib().oqpb_record_quartets(self._h, C.c_int(1 if enable else 0))

    def quartets(self):
        n = lib().oqpb_get_quartets(self._h, None, C.c_longlong(0))
        out = np.zeros((n, 4), dtype=np.int32)
        if n:
            lib().oqpb_get_quartets(self._h, _ip(out), C.c_longlong(n))
        return out

    def shell_density(self):
        ns = self.basis.nshell
        dsh = np.zeros((ns, ns))
        mx = C.c_double(0)
        self._check(lib().oqpb_get_shell_density(self._h, _dp(dsh), C.byref(mx)), "oqpb_get_shell_density")
        return dsh, mx.value

    def eri_block(self, i, j, k, l):
        out = np.zeros(10000)
        n = np.zeros(4, dtype=np.int32)
        self._check(lib().oqpb_eri_block(self._h, C.c_int(i), C.c_int(j), C.c_int(k), C.c_int(l), _dp(out), _ip(n)),
                    "oqpb_eri_block")
        return out[: int(np.prod(n))].reshape(tuple(int(x) for x in n)).copy()

    def rys(self, nroots, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        t2 = np.zeros((len(x), nroots))
        w = np.zeros((len(x), nroots))
        self._check(lib().oqpb_rys(self._h, C.c_int(nroots), C.c_int(len(x)), _dp(x), _dp(t2), _dp(w)), "oqpb_rys")
        return t2, w

    def fp64_peak_tflops(self) -> float:
        return lib().oqpb_fp64_peak_tflops(self._h)


# ------------------------------------------------------------------------------------------- consumers
class Int2RhfData:
    """int2_rhf_data_t (int2.F90:1414-1484): d, f packed (nfocks, ntri); f = raw accumulator unless post."""
    urohf = False

    def __init__(self, d, scale_exchange=1.0, scale_coulomb=1.0, post=False):
        self.d = np.ascontiguousarray(np.atleast_2d(d), dtype=np.float64)
        self.scale_exchange = scale_exchange
        self.scale_coulomb = scale_coulomb
        self.post = post
        self.f = None
        self.skipped = 0

    def _run(self, drv: Int2Compute):
        nf, ntri = self.d.shape
        assert ntri == drv.basis.ntri
        self.f = np.zeros_like(self.d)
        ns = C.c_longlong(0)
        drv._check(lib().oqpb_fock(drv._h, C.c_int(1 if self.urohf else 0), _dp(self.d), _dp(self.f), C.c_int(nf),
                                   C.c_double(self.scale_exchange), C.c_double(self.scale_coulomb),
                                   C.c_int(1 if self.post else 0), C.byref(ns)), "oqpb_fock")
        self.skipped = int(ns.value)


    def _run_cam(self, drv: Int2Compute, alpha, beta, mu, alpha_coulomb, beta_coulomb):
        nf, ntri = self.d.shape
        assert ntri == drv.basis.ntri
        self.f = np.zeros_like(self.d)
        ns = C.c_longlong(0)
        drv._check(lib().oqpb_fock_cam(drv._h, C.c_int(1 if self.urohf else 0), _dp(self.d), _dp(self.f), C.c_int(nf),
                                       C.c_double(alpha), C.c_double(beta), C.c_double(mu), C.c_double(alpha_coulomb),
                                       C.c_double(beta_coulomb), C.c_int(1 if self.post else 0), C.byref(ns)),
                   "oqpb_fock_cam")
        self.skipped = int(ns.value)


class Int2UrohfData(Int2RhfData):
    """int2_urohf_data_t (int2.F90:1488-1578): d(ntri,2) = alpha, beta."""
    urohf = True


class Int2TdData:
    """int2_td_data_t (tdhf_lib.F90:11-31).  d2: (nvec, nbf, nbf) with d2[v][mu, nu]; results apb, amb likewise
    (apb symmetrised as in parallel_stop, tdhf_lib.F90:107-109)."""

    def __init__(self, d2, int_apb=True, int_amb=False, tamm_dancoff=False, tamm_dancoff_coulomb=False,
                 scale_exchange=1.0, scale_coulomb=1.0):
        self.d2 = np.asarray(d2, dtype=np.float64)
        self.int_apb, self.int_amb = int_apb, int_amb
        self.tamm_dancoff, self.tamm_dancoff_coulomb = tamm_dancoff, tamm_dancoff_coulomb
        self.scale_exchange, self.scale_coulomb = scale_exchange, scale_coulomb
        self.apb = self.amb = None
        self.skipped = 0

    def _run(self, drv: Int2Compute):
        nv, nbf, _ = self.d2.shape
        dF = np.ascontiguousarray(np.transpose(self.d2, (0, 2, 1)))  # Fortran (mu, nu, v): mu fastest
        apb = np.zeros_like(dF)
        amb = np.zeros_like(dF)
        flags = ((OQPB_TD_APB if self.int_apb else 0) | (OQPB_TD_AMB if self.int_amb else 0) |
                 (OQPB_TD_TDA if self.tamm_dancoff else 0) | (OQPB_TD_TDA_COULOMB if self.tamm_dancoff_coulomb else 0))
        ns = C.c_longlong(0)
        drv._check(lib().oqpb_jk_td(drv._h, _dp(dF), C.c_int(nv), C.c_int(flags), C.c_double(self.scale_exchange),
                                    C.c_double(self.scale_coulomb), _dp(apb), _dp(amb), C.byref(ns)), "oqpb_jk_td")
        self.apb = np.transpose(apb, (0, 2, 1)).copy()
        self.amb = np.transpose(amb, (0, 2, 1)).copy()
        self.skipped = int(ns.value)

    def _run_cam(self, drv: Int2Compute, alpha, beta, mu, alpha_coulomb, beta_coulomb):
        nv, nbf, _ = self.d2.shape
        dF = np.ascontiguousarray(np.transpose(self.d2, (0, 2, 1)))
        apb = np.zeros_like(dF)
        amb = np.zeros_like(dF)
        flags = ((OQPB_TD_APB if self.int_apb else 0) | (OQPB_TD_AMB if self.int_amb else 0) |
                 (OQPB_TD_TDA if self.tamm_dancoff else 0) | (OQPB_TD_TDA_COULOMB if self.tamm_dancoff_coulomb else 0))
        ns = C.c_longlong(0)
        drv._check(lib().oqpb_jk_td_cam(drv._h, _dp(dF), C.c_int(nv), C.c_int(flags), C.c_double(alpha), C.c_double(beta),
                                        C.c_double(mu), C.c_double(alpha_coulomb), C.c_double(beta_coulomb), _dp(apb),
                                        _dp(amb), C.byref(ns)), "oqpb_jk_td_cam")
        self.apb = np.transpose(apb, (0, 2, 1)).copy()
        self.amb = np.transpose(amb, (0, 2, 1)).copy()
        self.skipped = int(ns.value)


class Int2MrsfData:
    """int2_mrsf_data_t (tdhf_mrsf_lib.F90:8-26).  d3: (nvec, ncomp, nbf, nbf) [v, c, mu, nu]; f3 likewise."""

    def __init__(self, d3, scale_exchange=1.0, scale_coulomb=1.0):
        self.d3 = np.asarray(d3, dtype=np.float64)
        self.scale_exchange, self.scale_coulomb = scale_exchange, scale_coulomb
        self.f3 = None
        self.skipped = 0

    def _run(self, drv: Int2Compute):
        nv, nc, nbf, _ = self.d3.shape
        dF = np.ascontiguousarray(np.transpose(self.d3, (3, 2, 1, 0)))  # Fortran d3(v, c, mu, nu): v fastest
        f3 = np.zeros_like(dF)
        ns = C.c_longlong(0)
        drv._check(lib().oqpb_jk_mrsf(drv._h, _dp(dF), C.c_int(nv), C.c_int(nc), C.c_double(self.scale_exchange),
                                      C.c_double(self.scale_coulomb), _dp(f3), C.byref(ns)), "oqpb_jk_mrsf")
        self.f3 = np.transpose(f3, (3, 2, 1, 0)).copy()
        self.skipped = int(ns.value)


def fock_jk(drv: Int2Compute, d, scale_exchange=1.0, scale_coulomb=1.0, urohf=False):
    """fock_jk (scf_addons.F90:1063-1214): packed d (nfocks, ntri) -> packed f, ready to use
    (0.5 scaling and diagonal doubling applied, :1177-1185); returns (f, nschwz)."""
    cons = (Int2UrohfData if urohf else Int2RhfData)(d, scale_exchange, scale_coulomb, post=True)
    drv.run(cons)
    return cons.f, cons.skipped


def _mrsf_run_cam(self, drv: Int2Compute, alpha, beta, mu, alpha_coulomb, beta_coulomb):
    """pass 1 all components (alpha_coulomb, alpha); pass 2 attenuated exchange of component 7 only with beta
    (tdhf_mrsf_lib.F90:312-326; beta_coulomb is not used by this consumer)"""
    nv, nc, nbf, _ = self.d3.shape
    dF = np.ascontiguousarray(np.transpose(self.d3, (3, 2, 1, 0)))
    f3 = np.zeros_like(dF)
    ns = C.c_longlong(0)
    drv._check(lib().oqpb_jk_mrsf_cam(drv._h, _dp(dF), C.c_int(nv), C.c_int(nc), C.c_double(alpha), C.c_double(beta),
                                      C.c_double(mu), C.c_double(alpha_coulomb), _dp(f3), C.byref(ns)), "oqpb_jk_mrsf_cam")
    self.f3 = np.transpose(f3, (3, 2, 1, 0)).copy()
    self.skipped = int(ns.value)


Int2MrsfData._run_cam = _mrsf_run_cam


class Int2TdgrdData:
    """int2_tdgrd_data_t (tdhf_lib.F90:33-36, update :228-295): the Z-vector / gradient response consumer with two spin
    blocks.  d2: (2, nbf, nbf) [s][mu, nu]; results apb (2, nbf, nbf) symmetrised, amb (2, nbf, nbf) (block 2 stays zero)."""

    def __init__(self, d2, int_apb=True, int_amb=False, scale_exchange=1.0, scale_coulomb=1.0):
        self.d2 = np.asarray(d2, dtype=np.float64)
        assert self.d2.shape[0] == 2
        self.int_apb, self.int_amb = int_apb, int_amb
        self.scale_exchange, self.scale_coulomb = scale_exchange, scale_coulomb
        self.apb = self.amb = None
        self.skipped = 0

    def _run(self, drv: Int2Compute):
        dF = np.ascontiguousarray(np.transpose(self.d2, (0, 2, 1)))  # Fortran (mu, nu, s)
        apb, amb = np.zeros_like(dF), np.zeros_like(dF)
        flags = (OQPB_TD_APB if self.int_apb else 0) | (OQPB_TD_AMB if self.int_amb else 0)
        ns = C.c_longlong(0)
        drv._check(lib().oqpb_jk_tdgrd(drv._h, _dp(dF), C.c_int(flags), C.c_double(self.scale_exchange),
                                       C.c_double(self.scale_coulomb), _dp(apb), _dp(amb), C.byref(ns)), "oqpb_jk_tdgrd")
        self.apb = np.transpose(apb, (0, 2, 1)).copy()
        self.amb = np.transpose(amb, (0, 2, 1)).copy()
        self.skipped = int(ns.value)


class Int2RpagrdData:
    """int2_rpagrd_data_t (tdhf_lib.F90:42-57, 1068-1320): xpy, xmy, t: (n, nspin, nbf, nbf) [q][s][mu, nu] or None;
    results hpp = H+[X+Y], hpt = H+[T] (symmetrised), hmm = H-[X-Y] with the shapes of xpy, t, xmy.
    X+Y and T must be symmetric matrices (as at every call site of the reference)."""

    def __init__(self, xpy=None, xmy=None, t=None, nspin=1, scale_exchange=1.0, scale_coulomb=1.0):
        self.nspin = nspin
        self.xpy, self.xmy, self.t = xpy, xmy, t
        self.scale_exchange, self.scale_coulomb = scale_exchange, scale_coulomb
        self.hpp = self.hpt = self.hmm = None
        self.skipped = 0

    def _run(self, drv: Int2Compute):
        nbf, nspin = drv.basis.nbf, self.nspin

        def prep(a):
            if a is None:
                return np.zeros((0, nspin, nbf, nbf)), 0
            a = np.asarray(a, dtype=np.float64)
            assert a.shape[1:] == (nspin, nbf, nbf)
            return np.ascontiguousarray(np.transpose(a, (0, 1, 3, 2))), a.shape[0]

        X, npp = prep(self.xpy)
        M, nm = prep(self.xmy)
        T, nt = prep(self.t)
        hpp, hpt, hmm = np.zeros_like(X), np.zeros_like(T), np.zeros_like(M)
        ns = C.c_longlong(0)
        drv._check(lib().oqpb_jk_rpagrd(drv._h, C.c_int(nspin), C.c_int(npp), C.c_int(nm), C.c_int(nt),
                                        _dp(X) if npp else None, _dp(M) if nm else None, _dp(T) if nt else None,
                                        C.c_double(self.scale_exchange), C.c_double(self.scale_coulomb),
                                        _dp(hpp) if npp else None, _dp(hpt) if nt else None, _dp(hmm) if nm else None,
                                        C.byref(ns)), "oqpb_jk_rpagrd")
        tr = lambda a: np.transpose(a, (0, 1, 3, 2)).copy()
        self.hpp, self.hpt, self.hmm = tr(hpp), tr(hpt), tr(hmm)
        self.skipped = int(ns.value)


class Int2UmrsfData:
    """int2_umrsf_data_t (tdhf_mrsf_lib.F90:28-32, update :337-426).  d3: (nvec, 11, nbf, nbf); f3 likewise."""

    def __init__(self, d3, scale_exchange=1.0, scale_coulomb=1.0):
        self.d3 = np.asarray(d3, dtype=np.float64)
        assert self.d3.shape[1] == 11
        self.scale_exchange, self.scale_coulomb = scale_exchange, scale_coulomb
        self.f3 = None
        self.skipped = 0

    def _run(self, drv: Int2Compute):
        nv = self.d3.shape[0]
        dF = np.ascontiguousarray(np.transpose(self.d3, (3, 2, 1, 0)))
        f3 = np.zeros_like(dF)
        ns = C.c_longlong(0)
        drv._check(lib().oqpb_jk_umrsf(drv._h, _dp(dF), C.c_int(nv), C.c_double(self.scale_exchange),
                                       C.c_double(self.scale_coulomb), _dp(f3), C.byref(ns)), "oqpb_jk_umrsf")
        self.f3 = np.transpose(f3, (3, 2, 1, 0)).copy()
        self.skipped = int(ns.value)

    def _run_cam(self, drv: Int2Compute, alpha, beta, mu, alpha_coulomb, beta_coulomb):
        nv = self.d3.shape[0]
        dF = np.ascontiguousarray(np.transpose(self.d3, (3, 2, 1, 0)))
        f3 = np.zeros_like(dF)
        ns = C.c_longlong(0)
        drv._check(lib().oqpb_jk_umrsf_cam(drv._h, _dp(dF), C.c_int(nv), C.c_double(alpha), C.c_double(beta), C.c_double(mu),
                                           C.c_double(alpha_coulomb), _dp(f3), C.byref(ns)), "oqpb_jk_umrsf_cam")
        self.f3 = np.transpose(f3, (3, 2, 1, 0)).copy()
        self.skipped = int(ns.value)


def jk(drv: Int2Compute, P, want_j=None, want_k=None):
    """oqpb_jk: generic J/K of n general matrices P (n, nbf, nbf) [m][mu, nu]:
    J_m(a,b) = sum_cd (ab|cd) P_m(c,d), K_m(a,c) = sum_bd (ab|cd) P_m(b,d).  Returns (J, K, nskipped)."""
    P = np.asarray(P, dtype=np.float64)
    n, nbf, _ = P.shape
    wj = np.ones(n, dtype=np.int32) if want_j is None else np.asarray(want_j, dtype=np.int32)
    wk = np.ones(n, dtype=np.int32) if want_k is None else np.asarray(want_k, dtype=np.int32)
    PF = np.ascontiguousarray(np.transpose(P, (0, 2, 1)))
    J, K = np.zeros_like(PF), np.zeros_like(PF)
    ns = C.c_longlong(0)
    drv._check(lib().oqpb_jk(drv._h, C.c_int(n), _dp(PF), _ip(wj), _ip(wk), _dp(J), _dp(K), C.byref(ns)), "oqpb_jk")
    return np.transpose(J, (0, 2, 1)).copy(), np.transpose(K, (0, 2, 1)).copy(), int(ns.value)


class RoutecSig:
    """Host-side mirror of the reference's `routec_sig` module (source/modules/routec_sig.F90:204-252): the MRSF Davidson
    sigma session  routec_sig_begin -> routec_sig_apply (per iteration) -> routec_sig_end  on the device.  `drv` must be
    initialised (basis, cutoff, screening); it is registered as the library's default context like the legacy seam."""

    def __init__(self, drv: Int2Compute):
        self.drv = drv
        self.active = False

    def begin(self, mo_a, mo_b, fa, fb, nocca, noccb, mrst, scale=1.0):
        """routec_sig_begin (routec_sig.F90:204-219): MO coefficients [mu, p], MO-basis Fock matrices, mrst 1 / 3.
        Returns the library's error code (0 = the session is ready)."""
        L = lib()
        L.oqpb_set_default_ctx(self.drv._h)
        nbf = self.drv.basis.nbf
        mats = [np.asfortranarray(np.asarray(m, dtype=np.float64)) for m in (mo_a, mo_b, fa, fb)]
        for m in mats:
            assert m.shape == (nbf, nbf)
        kind = 3 if mrst == 3 else 1
        rc = L.routec_sig_init(C.byref(C.c_int(nbf)), *[m.ctypes.data_as(C.POINTER(C.c_double)) for m in mats],
                               C.byref(C.c_int(nocca)), C.byref(C.c_int(noccb)), C.byref(C.c_int(kind)))
        if rc == 0:
            L.routec_sig_set_scale(C.byref(C.c_double(scale)))
            self.active = True
            self.ntrial = nocca * (nbf - noccb)
        return int(rc)

    def apply(self, bvec_mo):
        """routec_sig_apply (routec_sig.F90:221-241): bvec_mo (ntrial, nv) -> sigma (ntrial, nv) = (A-B) X, or None when the
        library declines (the reference then reverts to its native path)."""
        if not self.active:
            raise Int2Error("routec_sig: no session")
        b = np.asfortranarray(np.asarray(bvec_mo, dtype=np.float64))
        assert b.ndim == 2 and b.shape[0] == self.ntrial
        out = np.zeros_like(b, order="F")
        info = C.c_int(1)
        lib().routec_sig_iter(b.ctypes.data_as(C.POINTER(C.c_double)), C.byref(C.c_int(b.shape[1])),
                              out.ctypes.data_as(C.POINTER(C.c_double)), C.byref(info))
        return out if info.value == 0 else None

    def end(self):
        """routec_sig_end (routec_sig.F90:243-245)"""
        lib().routec_sig_free()
        self.active = False
