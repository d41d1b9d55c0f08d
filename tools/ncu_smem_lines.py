#!/usr/bin/env python3
"""Per-source-line shared-memory wavefronts / global sectors of an ncu report: python tools/ncu_smem_lines.py <rep> [top]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Line No')
h = rows[hi]; ci = {x: i for i, x in enumerate(h)}
W, E, G, I = ci['L1 Wavefronts Shared'], ci['L1 Wavefronts Shared Excessive'], ci['L2 Theoretical Sectors Global'], ci['Instructions Executed']
agg = collections.OrderedDict(); cur = None
for r in rows[hi + 1:]:
    if len(r) <= W: continue
    if r[0] != '':  # a CUDA source line header row
        cur = (r[0], r[1].strip()[:95]); agg.setdefault(cur, [0, 0, 0, 0]); continue
    if cur is None: continue
    try:
        a = agg[cur]; a[0] += int(r[W] or 0); a[1] += int(r[E] or 0); a[2] += int(r[G] or 0); a[3] += int(r[I] or 0)
    except ValueError:
        pass
tw = sum(a[0] for a in agg.values()); tg = sum(a[2] for a in agg.values()); ti = sum(a[3] for a in agg.values())
print(f"total smem wavefronts {tw:.3e} (excess {sum(a[1] for a in agg.values())/max(tw,1)*100:.0f}%)  global sectors {tg:.3e}  warp-instr {ti:.3e}")
for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*a[0]/max(tw,1):5.1f}% smem-wf (excess {100*a[1]/max(a[0],1):3.0f}%)  {100*a[2]/max(tg,1):5.1f}% gl-sect  {100*a[3]/max(ti,1):5.1f}% inst  L{ln:>4} {src}")
