#!/bin/bash
tag=${1:-s15}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
for wl in w32 c3 c2; do
timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_$wl.json 2> gpurun_out/${tag}_bench_$wl.err; python -c "import json; d=json.load(open('gpurun_out/${tag}_bench_$wl.json')); print('$wl', d['ms_per_step'], d['roofline']['frac'])"
done
