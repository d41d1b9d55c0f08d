#!/bin/bash
tag=${1:-s8}
mkdir -p gpurun_out
OQPB_KOWN=2 timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
OQPB_KOWN=2 timeout 600 python tools/class_profile.py w32 > gpurun_out/${tag}_class_w32_a60.txt 2>&1; head -1 gpurun_out/${tag}_class_w32_a60.txt
for v in a36 a20; do
OQPB_LIB=openqp_b200/libopenqp_b200_$v.so OQPB_KOWN=2 timeout 600 python tools/class_profile.py w32 > gpurun_out/${tag}_class_w32_$v.txt 2>&1; head -1 gpurun_out/${tag}_class_w32_$v.txt
done
