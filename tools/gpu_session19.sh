#!/bin/bash
tag=${1:-s19}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
for wl in c1 c2 c3 c5 w32; do
for g in 1 0; do
OQPB_GRAPH=$g timeout 600 python bench.py --workload $wl --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/${tag}_bench_${wl}_g$g.json 2> gpurun_out/${tag}_bench_${wl}_g$g.err; python -c "import json; d=json.load(open('gpurun_out/${tag}_bench_${wl}_g$g.json')); print('$wl graph=$g', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['roofline']['frac'])"
done
done
