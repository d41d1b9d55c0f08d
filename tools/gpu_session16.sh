#!/bin/bash
tag=${1:-s16}
mkdir -p gpurun_out
timeout 600 python tools/class_profile.py w32 > gpurun_out/${tag}_class_w32_base.txt 2>&1; head -1 gpurun_out/${tag}_class_w32_base.txt
for v in kp2 kp2r2; do
OQPB_LIB=openqp_b200/libopenqp_b200_$v.so timeout 600 python tools/class_profile.py w32 > gpurun_out/${tag}_class_w32_$v.txt 2>&1; head -1 gpurun_out/${tag}_class_w32_$v.txt
done
