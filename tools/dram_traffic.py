#!/usr/bin/env python3
"""HBM traffic of ONE Fock build (all ERI + enumeration launches), for bench.py's roofline.traffic.

  python tools/dram_traffic.py <workload>          (on the GPU box; writes profiles/r02_dram_<workload>.json)

Re-runs itself under `ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`
and brackets the third build of the workload with cudaProfilerStart/Stop, so that setup (Schwarz matrix, pair table) and the
warm-up builds are not counted.  ncu serialises the launches and replays each once per metric group: the byte counts are
per-launch sums, the durations are NOT a benchmark."""
import csv, ctypes, io, json, os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def child(wl):
    from openqp_b200 import workloads as W
    from openqp_b200.int2 import Int2Compute, Int2RhfData
    from openqp_b200.scf import pack
    mol, bs = W.build(wl)
    drv = Int2Compute(0).init(bs)
    drv.set_screening()
    d = pack(W.synthetic_density(bs))
    rt = ctypes.CDLL("libcudart.so")
    for it in range(3):
        if it == 2: rt.cudaProfilerStart()
        drv.run(Int2RhfData(d))
        if it == 2: rt.cudaProfilerStop()
    st = drv.last_stats()
    print("BUILD", json.dumps({"nquartets": st["nquartets"], "flops": st["flops"], "ntri": bs.ntri, "nbf": bs.nbf}))


def main():
    wl = sys.argv[1]
    if len(sys.argv) > 2 and sys.argv[2] == "--child":
        return child(wl)
    cmd = ["ncu", "--profile-from-start", "off", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum",
           "--clock-control", "none", "--cache-control", "none", "--csv", sys.executable, os.path.abspath(__file__), wl, "--child"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    rows = [l for l in r.stdout.splitlines() if l.startswith('"')]
    build = [l for l in r.stdout.splitlines() if l.startswith("BUILD")]
    rd = csv.DictReader(io.StringIO("\n".join(rows)))
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = {"read": 0.0, "write": 0.0}; n = 0; fam = {}
    for row in rd:
        name, metric = row["Kernel Name"], row["Metric Name"]
        if not metric.startswith("dram__bytes"): continue
        v = float(row["Metric Value"].replace(",", "")) * unit.get(row["Metric Unit"], 1.0)
        k = "read" if "read" in metric else "write"
        tot[k] += v
        f = name.split("<")[0].split("(")[0].replace("void ", "").replace("oqpb::", "").replace("(anonymous namespace)::", "")
        fam.setdefault(f, [0.0, 0])
        fam[f][0] += v; fam[f][1] += 1
        n += 1
    info = json.loads(build[0][6:]) if build else {}
    out = {"workload": wl, "dram_bytes_per_build": tot["read"] + tot["write"], "dram_read_bytes": tot["read"], "dram_write_bytes": tot["write"],
           "launches": n // 2, "by_kernel_family_bytes": {k: v[0] for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0])},
           "algorithmic_bytes": 16 * info.get("ntri", 0), "build": info,
           "how": "ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum over the third build "
                  "(cudaProfilerStart/Stop), summed over every launch of the build; serialised launches, --cache-control none (L2 stays warm between launches as in a real build)"}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "profiles", f"r02_dram_{wl}.json"), "w"), indent=1)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"r02_dram_{wl}.json"), "w"), indent=1)
    print(json.dumps(out)[:600])
    if n == 0: print(r.stderr[-2000:])


if __name__ == "__main__":
    main()
