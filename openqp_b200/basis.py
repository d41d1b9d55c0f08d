"""Basis-set container and synthetic molecules for the J/K Fock-build path.

`BasisSet` holds exactly the arrays of the reference's `basis_set`
(/root/reference/source/basis_tools.F90:32-84): one shell per contraction row, SP rows split,
primitive coefficients normalised by `normalize_primitives` only (basis_tools.F90:277-303,
`gauss_norm` :612-623) -- contracted functions are NOT renormalised (basis_api.F90:316-317).
All indices are 0-based here; the C ABI takes them 0-based as well (see include/oqp_b200.h).

Basis data come from `openqp_b200/data/basis.json`, extracted from the reference's
GAMESS-format `basis_sets/*.basis` (public Basis Set Exchange data) by tools/extract_basis.py.
Shell order per atom is file order (SURVEY Appendix A "Shell order caveat").
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass

import numpy as np

ANGSTROM_TO_BOHR = 1.0 / 0.52917721090299996  # pyoqp/oqp/utils/constants.py:2

_DATA = os.path.join(os.path.dirname(__file__), "data", "basis.json")
_DB = None

# ispher=auto rule (pyoqp/oqp/molecule/oqpdata.py:51-68): Pople sets Cartesian, cc-pVXZ/def2 pure
_SPHERICAL_DEFAULT = {"sto-3g": False, "3-21g": False, "6-31g": False, "6-31g(d)": False,
                      "cc-pvdz": True, "def2-svp": True, "cc-pvtz": True}
_ALIASES = {"6-31g*": "6-31g(d)"}


def _db():
    global _DB
    if _DB is None:
        with open(_DATA) as f:
            _DB = json.load(f)
    return _DB


def ncart(l: int) -> int:
    return (l + 1) * (l + 2) // 2


def gauss_norm(e: float, l: int) -> float:
    """basis_tools.F90:612-623"""
    norms = [1.0, 0.5, 0.75, 1.875, 6.5625, 29.53125, 162.421875]
    f = e * math.sqrt(e)
    return math.pi * math.sqrt(math.pi) * norms[l] / (f * e ** l)


@dataclass
class Molecule:
    Z: np.ndarray        # (natom,) int
    xyz: np.ndarray      # (natom,3) Bohr
    name: str = ""

    @property
    def natom(self):
        return len(self.Z)

    def nuclear_repulsion(self) -> float:
        e = 0.0
        for a in range(self.natom):
            for b in range(a):
                e += self.Z[a] * self.Z[b] / np.linalg.norm(self.xyz[a] - self.xyz[b])
        return float(e)


class BasisSet:
    """Arrays mirror basis_tools.F90:32-51 (`am, ncontr, g_offset, origin, ao_offset, naos,
    harmonic, ex, cc`) plus `centers` (= shell_centers, basis_tools.F90:1316-1326)."""

    def __init__(self, mol: Molecule, name: str, spherical: bool | None = None):
        name = _ALIASES.get(name.lower(), name.lower())
        db = _db()[name]
        self.name = name
        self.mol = mol
        if spherical is None:
            spherical = _SPHERICAL_DEFAULT[name]
        self.spherical = bool(spherical)          # HARMONIC_ACTIVE
        am, ncontr, g_offset, origin, harmonic, ex, cc = [], [], [], [], [], [], []
        for ia, z in enumerate(mol.Z):
            for sh in db[str(int(z))]:
                l = sh["l"]
                am.append(l)
                ncontr.append(len(sh["ex"]))
                g_offset.append(len(ex))
                origin.append(ia)
                harmonic.append(1 if self.spherical else 0)
                for e, c in zip(sh["ex"], sh["cc"]):
                    ex.append(e)
                    cc.append(c / math.sqrt(gauss_norm(2.0 * e, l)))
        self.am = np.array(am, dtype=np.int32)
        self.ncontr = np.array(ncontr, dtype=np.int32)
        self.g_offset = np.array(g_offset, dtype=np.int32)
        self.origin = np.array(origin, dtype=np.int32)
        self.harmonic = np.array(harmonic, dtype=np.int32)
        self.ex = np.array(ex, dtype=np.float64)
        self.cc = np.array(cc, dtype=np.float64)
        self.nshell = len(am)
        self.nprim = len(ex)
        naos = [(2 * l + 1) if (self.spherical and l >= 2) else ncart(l) for l in am]
        self.naos = np.array(naos, dtype=np.int32)
        self.ao_offset = np.concatenate([[0], np.cumsum(naos)[:-1]]).astype(np.int32)
        self.nbf = int(sum(naos))
        self.ntri = self.nbf * (self.nbf + 1) // 2
        self.centers = np.ascontiguousarray(mol.xyz[self.origin], dtype=np.float64)
        self.mxam = int(self.am.max())

    def describe(self) -> str:
        return (f"{self.mol.name}/{self.name} {'5d/7f' if self.spherical else 'cart'}: "
                f"{self.mol.natom} atoms, {self.nshell} shells, {self.nbf} bf")


# ----------------------------------------------------------------------------- molecules
_H2O_ANG = np.array([[0.000000000, 0.000000000, -0.041061554],
                     [-0.533194329, 0.533194329, -0.614469223],
                     [0.533194329, -0.533194329, -0.614469223]])  # examples/HF/H2O_RHF-HF_ENERGY.inp:6-9


def water() -> Molecule:
    return Molecule(np.array([8, 1, 1]), _H2O_ANG * ANGSTROM_TO_BOHR, "H2O")


def water_dimer() -> Molecule:
    """examples/other/h2o-2_rhf_cc-pvtz_hf.inp:3-8"""
    ang = np.array([[0.447604201, 0.612029479, -0.202683075],
                    [-0.449170155, 0.780027039, -0.457558173],
                    [0.980626115, 0.986327433, -0.890562004],
                    [0.437604201, 3.598029479, -1.202683075],
                    [-0.439170155, 3.780027039, -1.457558173],
                    [0.970626115, 3.986327433, -1.890562004]])
    return Molecule(np.array([8, 1, 1, 8, 1, 1]), ang * ANGSTROM_TO_BOHR, "(H2O)2")


def benzene() -> Molecule:
    """tools/scf-converger-ml/geometries/t0_benzene.xyz"""
    c = [[0.0, 1.396792, 0.0], [1.209657, 0.698396, 0.0], [1.209657, -0.698396, 0.0],
         [0.0, -1.396792, 0.0], [-1.209657, -0.698396, 0.0], [-1.209657, 0.698396, 0.0]]
    h = [[0.0, 2.479452, 0.0], [2.147078, 1.239726, 0.0], [2.147078, -1.239726, 0.0],
         [0.0, -2.479452, 0.0], [-2.147078, -1.239726, 0.0], [-2.147078, 1.239726, 0.0]]
    return Molecule(np.array([6] * 6 + [1] * 6), np.array(c + h) * ANGSTROM_TO_BOHR, "benzene")


def alkane(n: int = 20) -> Molecule:
    """all-trans n-C_nH_{2n+2}: r_CC 1.53 A, r_CH 1.09 A, tetrahedral angles, chain along x
    (SURVEY 8d, config 3)."""
    rcc, rch = 1.53, 1.09
    th = math.acos(-1.0 / 3.0)          # tetrahedral angle
    dx = rcc * math.sin(th / 2.0)
    dy = rcc * math.cos(th / 2.0)
    Z, xyz = [], []
    carbons = [np.array([i * dx, (i % 2) * dy, 0.0]) for i in range(n)]
    hy = rch * math.cos(th / 2.0)
    hz = rch * math.sin(th / 2.0)
    for i, c in enumerate(carbons):
        Z.append(6)
        xyz.append(c)
    for i, c in enumerate(carbons):
        sgn = -1.0 if i % 2 == 0 else 1.0
        for s in (+1.0, -1.0):
            Z.append(1)
            xyz.append(c + np.array([0.0, sgn * hy, s * hz]))
    # terminal hydrogens continue the zig-zag
    for i, d in ((0, -1.0), (n - 1, +1.0)):
        c = carbons[i]
        up = 1.0 if i % 2 == 0 else -1.0
        Z.append(1)
        xyz.append(c + np.array([d * rch * math.sin(th / 2.0), up * rch * math.cos(th / 2.0), 0.0]))
    return Molecule(np.array(Z), np.array(xyz) * ANGSTROM_TO_BOHR, f"n-C{n}H{2 * n + 2}")


def _random_rotation(rng) -> np.ndarray:
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    a, b, c, d = q
    return np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                     [2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)],
                     [2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d]])


def water_cluster(nx: int = 4, ny: int = 4, nz: int = 4, spacing: float = 3.1,
                  seed: int = 20261017) -> Molecule:
    """(H2O)_{nx*ny*nz} on a cubic lattice, each monomer = the C1 water geometry rotated by a
    seeded random SO(3) matrix about its oxygen."""
    rng = np.random.default_rng(seed)
    mono = _H2O_ANG - _H2O_ANG[0]
    Z, xyz = [], []
    for ix in range(nx):
        for iy in range(ny):
            for iz in range(nz):
                R = _random_rotation(rng)
                o = np.array([ix, iy, iz], dtype=float) * spacing
                for a, z in enumerate((8, 1, 1)):
                    Z.append(z)
                    xyz.append(o + R @ mono[a])
    return Molecule(np.array(Z), np.array(xyz) * ANGSTROM_TO_BOHR, f"(H2O){nx * ny * nz}")


def chromophore(nc: int = 20) -> Molecule:
    """Planar all-trans polyene-like C20 N1 O1 H18-style chromophore built analytically (config 5):
    a zig-zag sp2 chain C_nc capped by an NH2 (iminium-like) end and a C=O end, one H per inner C."""
    r, rh = 1.40, 1.08
    th = math.radians(120.0)
    dx = r * math.sin(th / 2.0)
    dy = r * math.cos(th / 2.0)
    Z, xyz = [], []
    heavy = [np.array([i * dx, (i % 2) * dy, 0.0]) for i in range(nc + 2)]
    for i, p in enumerate(heavy):
        Z.append(7 if i == 0 else (8 if i == nc + 1 else 6))
        xyz.append(p)
    for i in range(1, nc + 1):
        sgn = -1.0 if i % 2 == 0 else 1.0
        Z.append(1)
        xyz.append(heavy[i] + np.array([0.0, sgn * rh, 0.0]))
    # two H on N
    Z.append(1)
    xyz.append(heavy[0] + np.array([0.0, -1.01, 0.0]))
    Z.append(1)
    xyz.append(heavy[0] + np.array([-1.01 * math.sin(th / 2.0), 1.01 * math.cos(th / 2.0), 0.0]))
    return Molecule(np.array(Z), np.array(xyz) * ANGSTROM_TO_BOHR, f"C{nc}NOH{nc + 2}")


def build(config: str):
    """Named workloads of BASELINE.json:configs -> (Molecule, BasisSet)."""
    config = config.lower()
    if config in ("c1", "h2o"):
        m = water()
        return m, BasisSet(m, "6-31g(d)")
    if config in ("c2", "benzene"):
        m = benzene()
        return m, BasisSet(m, "cc-pvdz")
    if config in ("c3", "c20h42"):
        m = alkane(20)
        return m, BasisSet(m, "def2-svp")
    if config in ("c4", "w64"):
        m = water_cluster(4, 4, 4)
        return m, BasisSet(m, "cc-pvtz")
    if config == "w32":
        m = water_cluster(4, 4, 2)
        return m, BasisSet(m, "cc-pvtz")
    if config == "w8":
        m = water_cluster(2, 2, 2)
        return m, BasisSet(m, "cc-pvtz")
    if config == "w2":
        m = water_cluster(2, 1, 1)
        return m, BasisSet(m, "cc-pvtz")
    if config in ("c5", "chromophore"):
        m = chromophore(20)
        return m, BasisSet(m, "6-31g(d)")
    raise ValueError(config)
