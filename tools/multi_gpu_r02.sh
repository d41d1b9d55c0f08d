#!/bin/bash
mkdir -p gpurun_out
P=gpurun_out
port=29700
run() { wl=$1; n=$2; steps=$3; port=$((port+1));
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --workload $wl --steps $steps --warmup 3 --no-cpu-baseline > $P/r02_bench_${wl}_${n}gpu.json 2> $P/r02_bench_${wl}_${n}gpu.err
  python -c "
import json; d=json.load(open('$P/r02_bench_${wl}_${n}gpu.json')); print('$wl', $n, round(d['ms_per_step'],2), d['detail'].get('ranks'))"; }
run w32 8 5
run w32 4 4
run w32 2 3
run c4 8 3
