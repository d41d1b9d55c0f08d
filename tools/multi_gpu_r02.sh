#!/bin/bash
# 8-GPU box: strong scaling of the headline workload (bench.py under torchrun, one rank per GPU): tools/multi_gpu_r02.sh [N ...]
mkdir -p gpurun_out
P=gpurun_out
port=29700
run() { wl=$1; n=$2; steps=$3; port=$((port+1));
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --workload $wl --steps $steps --warmup 3 --no-cpu-baseline > $P/r02_bench_${wl}_${n}gpu.json 2> $P/r02_bench_${wl}_${n}gpu.err
  python -c "
import json; d=json.load(open('$P/r02_bench_${wl}_${n}gpu.json')); print('$wl', $n, round(d['ms_per_step'],2), d['detail'].get('ranks'))"; }
for n in ${@:-8}; do run w32 $n 5; done
