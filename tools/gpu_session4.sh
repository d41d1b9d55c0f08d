#!/bin/bash
tag=s4x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
b() {  # b <name> [env...]
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/${tag}_bench_$name.json 2> gpurun_out/${tag}_bench_$name.err
  python -c "import json,sys; d=json.load(open('gpurun_out/${tag}_bench_$name.json')); print('$name', d['ms_per_step'], d['roofline']['frac'])"
}
b default
b ctas3 OQPB_MED_CTAS=3
b ctas2 OQPB_MED_CTAS=2
b buckets1 OQPB_RUN_BUCKETS=1
b buckets3 OQPB_RUN_BUCKETS=3
timeout 600 python tools/class_profile.py w32 > gpurun_out/${tag}_class_w32.txt 2>&1; head -2 gpurun_out/${tag}_class_w32.txt
OQPB_MED_CTAS=3 timeout 600 python tools/class_profile.py w32 > gpurun_out/${tag}_class_w32_ctas3.txt 2>&1; head -2 gpurun_out/${tag}_class_w32_ctas3.txt
