"""Minimal RHF/UHF SCF driver used by tests and examples to turn a J/K builder into energies
(the reference's SCF, scf.F90, is OUT OF SCOPE and stays on the host unchanged; this is only the
~60-line harness SURVEY.md section 7 step 1 asks for so golden energies can pin the Fock builder).

`fock2e(d_packed[nfocks, ntri]) -> f_packed` is any J/K builder with fock_jk semantics
(scf_addons.F90:1063-1214): RHF F2e = J[D] - 1/2 K[D]; UHF F2e_s = J[Da+Db] - K[Ds].
Energy formula scf_addons.F90:2050-2062.
"""
from __future__ import annotations

import numpy as np


def pack(m: np.ndarray) -> np.ndarray:
    """Square symmetric -> packed lower triangle, ij = i(i+1)/2 + j (0-based), int2.F90:1436-1447."""
    n = m.shape[0]
    i, j = np.tril_indices(n)
    return np.ascontiguousarray(m[i, j])


def unpack(p: np.ndarray, n: int) -> np.ndarray:
    m = np.zeros((n, n))
    i, j = np.tril_indices(n)
    m[i, j] = p
    m[j, i] = p
    return m


def scf(nbf, S, H, enuc, fock2e, nalpha, nbeta=None, maxit=60, conv=1e-9, verbose=False):
    uhf = nbeta is not None and nbeta != nalpha
    nbeta = nalpha if nbeta is None else nbeta
    s, U = np.linalg.eigh(S)
    X = U @ np.diag(s ** -0.5) @ U.T

    def diag(F):
        e, C = np.linalg.eigh(X.T @ F @ X)
        return e, X @ C

    _, C = diag(H)
    if uhf:
        D = [C[:, :nalpha] @ C[:, :nalpha].T, C[:, :nbeta] @ C[:, :nbeta].T]
    else:
        D = [2.0 * C[:, :nalpha] @ C[:, :nalpha].T]
    errs, focks = [], []
    e_old = 0.0
    for it in range(maxit):
        dp = np.stack([pack(d) for d in D])
        f2 = fock2e(dp)
        F = [H + unpack(f2[k], nbf) for k in range(len(D))]
        e = enuc + 0.5 * sum(np.sum(D[k] * (H + F[k])) for k in range(len(D)))
        err = np.concatenate([(F[k] @ D[k] @ S - S @ D[k] @ F[k]).ravel() for k in range(len(D))])
        if verbose:
            print(f"it {it:3d} E = {e:.12f} err = {np.abs(err).max():.2e}")
        if abs(e - e_old) < conv and np.abs(err).max() < 1e-7:
            return e, D, F
        e_old = e
        errs.append(err)
        focks.append(np.stack(F))
        errs, focks = errs[-8:], focks[-8:]
        if len(errs) > 1:
            n = len(errs)
            B = -np.ones((n + 1, n + 1))
            B[n, n] = 0
            for a in range(n):
                for b in range(n):
                    B[a, b] = errs[a] @ errs[b]
            rhs = np.zeros(n + 1)
            rhs[n] = -1
            try:
                c = np.linalg.solve(B, rhs)[:n]
                Fd = sum(c[a] * focks[a] for a in range(n))
            except np.linalg.LinAlgError:
                Fd = focks[-1]
        else:
            Fd = focks[-1]
        if uhf:
            Ca, Cb = diag(Fd[0])[1], diag(Fd[1])[1]
            D = [Ca[:, :nalpha] @ Ca[:, :nalpha].T, Cb[:, :nbeta] @ Cb[:, :nbeta].T]
        else:
            C = diag(Fd[0])[1]
            D = [2.0 * C[:, :nalpha] @ C[:, :nalpha].T]
    raise RuntimeError("SCF not converged")
