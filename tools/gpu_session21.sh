#!/bin/bash
tag=${1:-s21}
mkdir -p gpurun_out
OQPB_KOWN=2 OQPB_LIB=openqp_b200/libopenqp_b200_a60r.so timeout 600 python tools/class_profile.py w32 > gpurun_out/${tag}_class_w32_a60r.txt 2>&1; head -1 gpurun_out/${tag}_class_w32_a60r.txt
for wl in c1 c2 c3; do
timeout 600 python bench.py --workload $wl --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/${tag}_bench_${wl}.json 2> gpurun_out/${tag}_bench_${wl}.err; python -c "import json; d=json.load(open('gpurun_out/${tag}_bench_${wl}.json')); print('$wl', round(d['ms_per_step'],3), d['roofline']['frac'])"
done
