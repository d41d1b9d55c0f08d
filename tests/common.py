"""Shared helpers for the parity tests."""
import json
import os

import numpy as np

from openqp_b200 import basis as B
from openqp_b200.scf import pack, scf, unpack

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def random_sym_density(nbf, seed=1234, scale=1.0):
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(nbf, nbf)) * scale
    return d + d.T


def decaying_density(bs, seed=7):
    """Density-like symmetric matrix whose elements decay with inter-shell distance, so that the
    Schwarz x density screening actually removes quartets (unlike a dense random matrix)."""
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(bs.nbf, bs.nbf))
    d = d + d.T
    cen = np.repeat(bs.centers, bs.naos, axis=0)
    r = np.linalg.norm(cen[:, None, :] - cen[None, :, :], axis=2)
    return d * np.exp(-0.6 * r)


def rpa_energies(bs, S, H, enuc, nocc, fock2e, td_apb_amb, nroots=3):
    """RHF ground state then singlet RPA: (A-B)(A+B) Z = w^2 Z with
    (A+B) X = de X + Co^T apb[P] Cv,  (A-B) X = de X + Co^T amb[P] Cv,  P = Co X Cv^T
    (consumer semantics tdhf_lib.F90:140-224; driver modules/tdhf_energy.F90:214-258)."""
    e, D, F = scf(bs.nbf, S, H, enuc, fock2e, nocc, conv=1e-11)
    s, U = np.linalg.eigh(S)
    X = U @ np.diag(s ** -0.5) @ U.T
    eps, C = np.linalg.eigh(X.T @ F[0] @ X)
    C = X @ C
    Co, Cv = C[:, :nocc], C[:, nocc:]
    nv = Cv.shape[1]
    P = np.zeros((nocc * nv, bs.nbf, bs.nbf))
    for i in range(nocc):
        for a in range(nv):
            P[i * nv + a] = np.outer(Co[:, i], Cv[:, a])
    apb, amb = td_apb_amb(P)
    de = (eps[nocc:][None, :] - eps[:nocc][:, None]).ravel()
    ApB = np.array([(Co.T @ apb[k] @ Cv).ravel() for k in range(nocc * nv)]) + np.diag(de)
    AmB = np.array([(Co.T @ amb[k] @ Cv).ravel() for k in range(nocc * nv)]) + np.diag(de)
    w2 = np.linalg.eigvals(AmB @ ApB)
    w = np.sort(np.sqrt(np.real(w2)))
    return e, w[:nroots]
