#!/bin/bash
tag=${1:-s30}
mkdir -p gpurun_out
P=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $P/${tag}_tests.log 2>&1; tail -2 $P/${tag}_tests.log
b() { name=$1; shift; timeout 900 python bench.py "$@" > $P/r02_bench_$name.json 2> $P/r02_bench_$name.err; python -c "
import json; d=json.load(open('$P/r02_bench_$name.json')); print('$name', round(d['ms_per_step'],2), 'ms', d.get('roofline',{}).get('frac'), d.get('cpu_baseline',{}).get('value'))"; }
b w32 --steps 5 --warmup 3
b w32_cam --cam --steps 3 --warmup 3 --no-cpu-baseline
b c1 --workload c1 --steps 20 --warmup 5
b c2 --workload c2 --steps 20 --warmup 5
b c3 --workload c3 --steps 10 --warmup 3
b c5 --workload c5 --steps 3 --warmup 3
b c5_cut1e-8 --workload c5 --cutoff 1e-8 --steps 3 --warmup 3 --no-cpu-baseline
timeout 600 python tools/class_profile.py w32 > $P/r02_class_profile_w32.txt 2>&1; head -1 $P/r02_class_profile_w32.txt
