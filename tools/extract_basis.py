#!/usr/bin/env python3
"""Extract the element blocks the benchmark/test molecules need from the GAMESS-format
basis files shipped with the reference (public Basis Set Exchange data) into one compact
JSON fixture that travels with the repo (the GPU box has no /root/reference).

Run in the build container only:  python tools/extract_basis.py
Format parsed: `basis_sets/*.basis` ($DATA, ELEMENT name, `L nprim`, `idx exp coef [coef_p]`);
L-shells ("L"=SP) are split into one S and one P row exactly like the reference's loader
(pyoqp/oqp/library/set_basis.py:133-153 -> one shell per contraction row, zero coefficients dropped).
"""
import json, os, sys

REF = "/root/reference/basis_sets"
FILES = {"sto-3g": "sto-3g.basis", "3-21g": "3-21g.basis", "6-31g": "6-31g.basis",
         "6-31g(d)": "6-31g(d).basis", "cc-pvdz": "cc-pvdz.basis", "def2-svp": "def2-svp.basis",
         "cc-pvtz": "cc-pvtz.basis"}
ELEMENTS = {"HYDROGEN": 1, "CARBON": 6, "NITROGEN": 7, "OXYGEN": 8}
LMAP = {"S": 0, "P": 1, "D": 2, "F": 3, "G": 4, "H": 5, "I": 6}


def parse(path):
    out = {}
    lines = open(path).read().splitlines()
    i = 0
    while i < len(lines) and not lines[i].strip().upper().startswith("$DATA"):
        i += 1
    i += 1
    cur = None
    while i < len(lines):
        s = lines[i].strip()
        i += 1
        if not s or s.startswith("!"):
            continue
        if s.upper().startswith("$END"):
            break
        tok = s.split()
        if len(tok) == 1 and tok[0].isalpha():
            cur = tok[0].upper()
            out[cur] = []
            continue
        if tok[0].upper() in LMAP or tok[0].upper() == "L":
            typ, n = tok[0].upper(), int(tok[1])
            ex, c1, c2 = [], [], []
            for _ in range(n):
                t = lines[i].split()
                i += 1
                ex.append(float(t[1].replace("D", "E")))
                c1.append(float(t[2].replace("D", "E")))
                if typ == "L":
                    c2.append(float(t[3].replace("D", "E")))
            if typ == "L":
                out[cur].append({"l": 0, "ex": ex, "cc": c1})
                out[cur].append({"l": 1, "ex": ex, "cc": c2})
            else:
                out[cur].append({"l": LMAP[typ], "ex": ex, "cc": c1})
    return out


def main():
    db = {}
    for name, fn in FILES.items():
        allel = parse(os.path.join(REF, fn))
        db[name] = {}
        for el, z in ELEMENTS.items():
            if el not in allel:
                continue
            shells = []
            for sh in allel[el]:
                keep = [(e, c) for e, c in zip(sh["ex"], sh["cc"]) if c != 0.0]
                shells.append({"l": sh["l"], "ex": [k[0] for k in keep], "cc": [k[1] for k in keep]})
            db[name][str(z)] = shells
    dst = os.path.join(os.path.dirname(__file__), "..", "openqp_b200", "data", "basis.json")
    json.dump(db, open(dst, "w"), indent=0, separators=(",", ":"))
    for name in db:
        print(name, {z: [(s["l"], len(s["ex"])) for s in sh] for z, sh in db[name].items()})


if __name__ == "__main__":
    main()
