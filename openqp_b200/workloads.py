"""Synthetic inputs for the benchmark configurations (BASELINE.json:configs).

Densities are synthetic (no SCF is run here): a symmetric matrix with O(1) intra-atomic blocks whose
inter-shell elements decay exponentially with distance, which reproduces the locality that makes the
Schwarz x density screening (int2.F90:975-986) remove most far quartets.  Seeds are fixed.
"""
from __future__ import annotations

import numpy as np

from . import basis as B

WORKLOADS = {
    "c1": "H2O RHF/6-31G(d)",
    "c2": "benzene B3LYP/cc-pVDZ (J/K part, scale_exchange=0.2)",
    "c3": "n-C20H42 RHF/def2-SVP",
    "w32": "(H2O)32 RHF/cc-pVTZ (96 atoms, 1856 bf)",
    "c4": "(H2O)64 RHF/cc-pVTZ (192 atoms, 3712 bf)",
    "w8": "(H2O)8 RHF/cc-pVTZ",
    "c5": "C20NOH22 chromophore MRSF/6-31G(d) (nvec x 7 densities)",
}


def synthetic_density(bs, seed: int = 7, decay: float = 0.6, scale: float = 1.0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(bs.nbf, bs.nbf))
    d = d + d.T
    cen = np.repeat(bs.centers, bs.naos, axis=0)
    # pairwise distances without an (n,n,3) temporary
    g = cen @ cen.T
    sq = np.diag(g)
    r = np.sqrt(np.maximum(sq[:, None] + sq[None, :] - 2 * g, 0.0))
    return scale * d * np.exp(-decay * r)


def scale_exchange(name: str) -> float:
    return 0.2 if name == "c2" else 1.0


def build(name: str):
    mol, bs = B.build(name)
    return mol, bs


def mrsf_densities(bs, nvec: int, ncomp: int = 7, seed: int = 7, decay: float = 0.6) -> np.ndarray:
    """d3(nvec, 7, nbf, nbf): general (non-symmetric) trial densities of the shape `mrsfcbc` produces
    (tdhf_mrsf_lib.F90:940-1010), here random with the same distance decay as `synthetic_density`."""
    rng = np.random.default_rng(seed)
    cen = np.repeat(bs.centers, bs.naos, axis=0)
    g = cen @ cen.T
    sq = np.diag(g)
    r = np.sqrt(np.maximum(sq[:, None] + sq[None, :] - 2 * g, 0.0))
    env = np.exp(-decay * r)
    return rng.normal(size=(nvec, ncomp, bs.nbf, bs.nbf)) * env[None, None] * 0.1
