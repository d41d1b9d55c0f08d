#!/bin/bash
tag=${1:-s20}
mkdir -p gpurun_out
for wl in c1 c2 c3; do
for l in 4 8; do
OQPB_NLANES=$l timeout 600 python bench.py --workload $wl --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/${tag}_bench_${wl}_l$l.json 2> gpurun_out/${tag}_bench_${wl}_l$l.err; python -c "import json; d=json.load(open('gpurun_out/${tag}_bench_${wl}_l$l.json')); print('$wl lanes=$l', round(d['ms_per_step'],3), d['roofline']['frac'])"
done
done
