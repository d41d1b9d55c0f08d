#!/bin/bash
# 8-GPU box: strong scaling of w32 and c4 (bench.py under torchrun, one rank per GPU) + the single-process multi-device context
mkdir -p gpurun_out
P=gpurun_out
nvidia-smi -L | wc -l
port=29600
run() { wl=$1; n=$2; steps=$3; port=$((port+1));
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --workload $wl --steps $steps --warmup 3 --no-cpu-baseline > $P/r02_bench_${wl}_${n}gpu.json 2> $P/r02_bench_${wl}_${n}gpu.err
  python -c "
import json; d=json.load(open('$P/r02_bench_${wl}_${n}gpu.json')); print('$wl', $n, round(d['ms_per_step'],2), d['detail'].get('ranks'))"; }
run w32 8 5
run w32 4 5
run w32 2 5
run c4 8 3
run c4 4 3
run c4 2 2
OQPB_WHOLE_MS=0 run w32 8 5 && mv $P/r02_bench_w32_8gpu.json $P/r02_bench_w32_8gpu_split_only.json
run w32 8 5
timeout 600 python - > $P/r02_multi_ctx_w32.txt 2>&1 <<'PY'
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from openqp_b200 import workloads as W
from openqp_b200.int2 import Int2Compute, Int2RhfData
from openqp_b200.scf import pack
mol, bs = W.build("w32")
d = pack(W.synthetic_density(bs))
ref = None
for nd in (1, 8):
    drv = Int2Compute(0, ndevices=nd).init(bs); drv.set_screening()
    for r in range(3):
        t = time.perf_counter(); c = drv.run(Int2RhfData(d)); dt = time.perf_counter() - t
    if ref is None: ref = c.f.copy()
    print(f"oqpb_ctx_create_multi({nd}) w32: oqpb_fock with host buffers {1e3*dt:.1f} ms per build, max|F - F_1gpu| = {np.abs(c.f - ref).max():.2e}", flush=True)
    drv.clean()
PY
cat $P/r02_multi_ctx_w32.txt
