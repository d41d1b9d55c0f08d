#!/bin/bash
# runtime knob sweep on the GPU box: build time of the last of 3 builds per setting
wl=${1:-w32}
for s in "" "OQPB_NLANES=2" "OQPB_NLANES=8" "OQPB_TASK_CAP_LOG2=21" "OQPB_TASK_CAP_LOG2=25" "OQPB_GRID_PCT=50" "OQPB_GRID_PCT=200" "OQPB_TASK_CAP_LOG2=25 OQPB_NLANES=8"; do
  echo "== $s: $(env $s python tools/run_build.py $wl 3 | tail -1)"
done
