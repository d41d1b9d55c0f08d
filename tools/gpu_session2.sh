#!/bin/bash
# session 2: tests + run-kernel variants
tag=s2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
b() {  # b <name> [env...]
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/${tag}_bench_$name.json 2> gpurun_out/${tag}_bench_$name.err
  python -c "import json,sys; d=json.load(open('gpurun_out/${tag}_bench_$name.json')); print('$name', d['ms_per_step'], d['roofline']['frac'])"
}
b norun OQPB_RUN=0
b run
b k0 OQPB_LIB=$PWD/openqp_b200/libopenqp_b200_k0.so
b med OQPB_LIB=$PWD/openqp_b200/libopenqp_b200_med.so
timeout 600 python tools/class_profile.py w32 > gpurun_out/${tag}_class_w32_run.txt 2>&1; head -2 gpurun_out/${tag}_class_w32_run.txt
OQPB_LIB=$PWD/openqp_b200/libopenqp_b200_med.so timeout 600 python tools/class_profile.py w32 > gpurun_out/${tag}_class_w32_med.txt 2>&1; head -2 gpurun_out/${tag}_class_w32_med.txt
