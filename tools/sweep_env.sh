#!/bin/bash
# runtime-knob sweep on one GPU: tools/sweep_env.sh   (kernel_ms of the third w32 build per setting)
run() { name=$1; shift; env "$@" timeout 300 python tools/run_build.py w32 3 2>&1 | tail -1 | sed "s/^/$name: /"; }
run default X=1
run grid70 OQPB_GRID_PCT=70
run grid40 OQPB_GRID_PCT=40
run cap26 OQPB_TASK_CAP_LOG2=26
for wl in c2 c3; do timeout 300 python bench.py --workload $wl --steps 10 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl', round(d['ms_per_step'],3))"; done
