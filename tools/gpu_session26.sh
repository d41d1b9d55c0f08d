#!/bin/bash
tag=${1:-s26}
mkdir -p gpurun_out
timeout 600 python tools/class_profile.py w32 > gpurun_out/${tag}_class_w32_def.txt 2>&1; head -1 gpurun_out/${tag}_class_w32_def.txt
for v in capa capb; do
OQPB_LIB=openqp_b200/libopenqp_b200_$v.so timeout 600 python tools/class_profile.py w32 > gpurun_out/${tag}_class_w32_$v.txt 2>&1; head -1 gpurun_out/${tag}_class_w32_$v.txt
done
