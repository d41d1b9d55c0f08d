#!/bin/bash
tag=${1:-s10}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; python -c "import json; d=json.load(open('gpurun_out/${tag}_bench.json')); print(d['ms_per_step'], d['roofline']['frac'])"
OQPB_KOWN=2 timeout 600 python tools/class_profile.py w32 > gpurun_out/${tag}_class_w32_k2.txt 2>&1; head -1 gpurun_out/${tag}_class_w32_k2.txt
OQPB_LIB=openqp_b200/libopenqp_b200_r96.so OQPB_KOWN=2 timeout 600 python tools/class_profile.py w32 > gpurun_out/${tag}_class_w32_r96.txt 2>&1; head -1 gpurun_out/${tag}_class_w32_r96.txt
