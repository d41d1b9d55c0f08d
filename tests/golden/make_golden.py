#!/usr/bin/env python3
"""Regenerate tests/golden/*.json.  Run in the BUILD container only (reads /root/reference):
    python tests/golden/make_golden.py
 - reference_energies.json : pure-HF golden energies copied from the reference's example decks
 - reference_pure_tables.json : l=2..4 Cartesian->pure projection coefficients parsed from
   source/integrals/int2_pure_generated.F90 (load_l2 / load_l3 / load_l4)
 - oracle_fock_h2o.json : oracle Fock vectors for fixed seeded densities (H2O 6-31G(d) and cc-pVTZ) so the GPU
   tests also compare against committed vectors, not only against a live oracle build
"""
import json, os, re, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def energies():
    out = {}
    for key, path in {
        "h2o_rhf_631gd": "examples/HF/H2O_RHF-HF_ENERGY.json",
        "h2o_uhf_triplet_631gd": "examples/HF/H2O_UHF-HF_ENERGY.json",
        "h2o_rohf_631gd": "examples/HF/H2O_ROHF-HF_ENERGY.json",
        "h2o_rhf_631g": "examples/other/h2o_rhf_6-31g_hf.json",
        "h2o_dimer_rhf_631gd": "examples/other/h2o-2_rhf_cc-pvtz_hf.json",
        "h2o_rhf_sto3g_nmr": "examples/NMR/H2O_RHF-NMR.json",
    }.items():
        d = json.load(open(os.path.join(REF, path)))
        out[key] = {"energy": d["energy"], "source": path}
    d = json.load(open(os.path.join(REF, "examples/TDHF/H2O_TDHF_ENERGY.json")))
    out["h2o_tdhf_631gd"] = {"energy": d["energy"], "td_energies": d["td_energies"], "source": "examples/TDHF/H2O_TDHF_ENERGY.json"}
    # MRSF-CIS (no functional) on an ROHF triplet: the only XC-free MRSF deck; pins int2_mrsf_data_t and the sigma triple.
    # roots_nstate20: the tighter-converged roots quoted in the deck's own header comment (nstate=20 reference run)
    path = "examples/MRSF-TDDFT/CH2O_MRSFTDDFT_SYMMETRY_BLOCK_COVERAGE.json"
    d = json.load(open(os.path.join(REF, path)))
    out["ch2o_mrsf_triplet_631g"] = {"energy": d["energy"], "td_energies": d["td_energies"], "atoms": d["atoms"], "coord": d["coord"],
                                     "roots_nstate20": [-0.00500108, 0.07496767, 0.17574606, 0.23436551], "source": path}
    json.dump(out, open(os.path.join(HERE, "reference_energies.json"), "w"), indent=1)


def pure_tables():
    src = open(os.path.join(REF, "source/integrals/int2_pure_generated.F90")).read()
    out = {}
    for l in (2, 3, 4):
        body = src[src.index(f"subroutine load_l{l}(proj)"):src.index(f"end subroutine load_l{l}")]
        rows = re.findall(r"add_term\(proj,\s*(\d+),\s*(\d+),\s*([-+0-9.eE]+)_dp\)", body)
        out[str(l)] = [[int(a), int(b), float(c)] for a, b, c in rows]
    json.dump(out, open(os.path.join(HERE, "reference_pure_tables.json"), "w"))


def oracle_fock():
    from openqp_b200 import basis as B
    from openqp_b200.scf import pack
    from oracle.oracle import Oracle
    out = {}
    for name in ("6-31g(d)", "cc-pvtz"):
        mol = B.water()
        bs = B.BasisSet(mol, name)
        o = Oracle(bs)
        q = o.set_screening()
        rng = np.random.default_rng(1234)
        d = rng.normal(size=(bs.nbf, bs.nbf))
        d = d + d.T
        f, st = o.fock(pack(d))
        out[name] = {"seed": 1234, "nbf": bs.nbf, "fock": f[0].tolist(), "stats": st,
                     "schwarz_sum": float(q.sum()), "schwarz_max": float(q.max())}
    json.dump(out, open(os.path.join(HERE, "oracle_fock_h2o.json"), "w"))


if __name__ == "__main__":
    energies()
    pure_tables()
    oracle_fock()
    print("ok")
