"""CPU: host-side logic -- basis layout of the five configs, product never imports the oracle,
and the N>1 replicated-data split + all-reduce (gloo, world_size 2) reassembles the full Fock matrix."""
import os
import subprocess
import sys

import numpy as np
import pytest

from openqp_b200 import basis as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config_sizes_match_survey():
    assert (B.build("c1")[1].nshell, B.build("c1")[1].nbf) == (10, 19)
    assert (B.build("c2")[1].nshell, B.build("c2")[1].nbf) == (54, 114)
    assert (B.build("c3")[1].nshell, B.build("c3")[1].nbf) == (246, 490)
    m, bs = B.build("c4")
    assert (m.natom, bs.nshell, bs.nbf, bs.ntri) == (192, 1408, 3712, 6891328)
    m, bs = B.build("c5")
    assert m.natom == 44 and bs.spherical is False


def test_geometries_sane():
    for cfg in ("c3", "c5", "w8"):
        m, _ = B.build(cfg)
        d = np.linalg.norm(m.xyz[:, None] - m.xyz[None], axis=2) + np.eye(m.natom) * 10
        assert d.min() > 1.6, (cfg, d.min())  # Bohr: no overlapping atoms


def test_primitive_normalisation():
    """normalize_primitives (basis_tools.F90:277-303): an uncontracted s primitive is unit-normalised."""
    bs = B.BasisSet(B.water(), "6-31g")
    s1 = [i for i in range(bs.nshell) if bs.am[i] == 0 and bs.ncontr[i] == 1][0]
    a, c = bs.ex[bs.g_offset[s1]], bs.cc[bs.g_offset[s1]]
    assert abs(c * c * (np.pi / (2 * a)) ** 1.5 - 1.0) < 1e-12


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "openqp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt.replace("oracle/ ", ""), f


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from openqp_b200 import basis as B
from openqp_b200.scf import pack
from oracle.oracle import Oracle
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
bs = B.BasisSet(B.water(), "6-31g(d)")
o = Oracle(bs); o.set_screening()
rng = np.random.default_rng(11); d = rng.normal(size=(bs.nbf, bs.nbf)); d = pack(d + d.T)
# replicated-data split of the bra pair list, int2.F90:759-761; one all-reduce of the partial Fock, :1396
f, st = o.fock(d, post=False, stride=2, offset=rank, nthreads=1)
t = torch.from_numpy(f.copy()); dist.all_reduce(t)
nq = torch.tensor([st["nquartets"]]); dist.all_reduce(nq)
full, st_full = o.fock(d, post=False, nthreads=1)
if rank == 0:
    assert np.abs(t.numpy() - full).max() < 1e-12, np.abs(t.numpy() - full).max()
    assert int(nq) == st_full["nquartets"]
    print("OK")
dist.destroy_process_group()
"""


def test_two_rank_split_and_allreduce_gloo(tmp_path):
    port = 29500 + os.getpid() % 2000
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "OK" in outs[0][0]


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU oracle timed on the host cores) prints one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for extra in ([], ["--cam"]):
        out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1",
                              "--warmup", "0", "--cpu-seconds", "0.5"] + extra, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-500:]
        lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
        assert len(lines) == 1
        d = json.loads(lines[0])
        for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
            assert k in d, k
        assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
        assert d["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in d["config"]
