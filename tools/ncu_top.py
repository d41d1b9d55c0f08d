#!/usr/bin/env python3
"""usage: tools/ncu_top.py <workload> <out-name> <kernel-regex> [<out-name> <kernel-regex> ...]
Two passes on the GPU box: (1) launch list with gpu__time_duration for the matching kernels, (2) one
`ncu --set full` capture of the LONGEST matching launch (--launch-skip).  Output: gpurun_out/<out-name>.ncu-rep"""
import csv, io, os, subprocess, sys
TARGET = os.environ.get("OQPB_NCU_TARGET", "tools/run_build.py")  # e.g. tools/mrsf_bench.py
wl = sys.argv[1]
pairs = list(zip(sys.argv[2::2], sys.argv[3::2]))
for out, rx in pairs:
    if "@" in out:  # <out-name>@<N>: capture the N-th matching launch, no duration pass
        out, best = out.split("@"); best = int(best)
        cmd = ["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "--kernel-name-base", "demangled",
               "-k", f"regex:{rx}", "--launch-skip", str(best), "-c", "1", "-o", f"gpurun_out/{out}", "-f",
               "python", TARGET, wl] + os.environ.get("OQPB_NCU_ARGS", "1").split()
        r = subprocess.run(cmd, capture_output=True, text=True)
        print(out, r.stdout[-200:], flush=True)
        continue
    cmd = ["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "--kernel-name-base", "demangled",
           "-k", f"regex:{rx}", "--csv", "python", TARGET, wl] + os.environ.get("OQPB_NCU_ARGS", "1").split()
    r = subprocess.run(cmd, capture_output=True, text=True)
    rows = [l for l in r.stdout.splitlines() if l.startswith('"')]
    rd = list(csv.DictReader(io.StringIO("\n".join(rows))))
    durs = []
    for row in rd:
        if row.get("Metric Name") == "gpu__time_duration.sum":
            v = float(row["Metric Value"].replace(",", ""))
            if row.get("Metric Unit") == "us": v *= 1e3
            elif row.get("Metric Unit") == "ms": v *= 1e6
            durs.append(v)
    if not durs:
        print(out, "no launches matched", r.stderr[-500:]); continue
    best = max(range(len(durs)), key=lambda i: durs[i])
    print(f"{out}: {len(durs)} launches, total {sum(durs)/1e6:.2f} ms, longest #{best} = {durs[best]/1e6:.3f} ms", flush=True)
    cmd = ["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "--kernel-name-base", "demangled",
           "-k", f"regex:{rx}", "--launch-skip", str(best), "-c", "1", "-o", f"gpurun_out/{out}", "-f",
           "python", TARGET, wl] + os.environ.get("OQPB_NCU_ARGS", "1").split()
    r = subprocess.run(cmd, capture_output=True, text=True)
    print(r.stdout[-300:])
