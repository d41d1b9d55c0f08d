#!/bin/bash
tag=${1:-s14}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "partition or multi" > gpurun_out/${tag}_tests_part.log 2>&1; tail -3 gpurun_out/${tag}_tests_part.log
for w in 2.0 3.0 4.0 6.0; do
echo "OQPB_WHOLE_MS=$w"
OQPB_WHOLE_MS=$w timeout 900 python tools/partition_bench.py w32 8 2>&1 | tail -1
done > gpurun_out/${tag}_partition_whole.txt 2>&1
cat gpurun_out/${tag}_partition_whole.txt
