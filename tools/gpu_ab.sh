#!/bin/bash
# A/B of tuning variants: tools/gpu_ab.sh <tag> <variant> [<variant> ...]  (class profiles of w32 with the default library and
# with openqp_b200/libopenqp_b200_<variant>.so, built by tools/tune_variants.sh)
tag=$1; shift
mkdir -p gpurun_out
timeout 600 python tools/class_profile.py w32 > gpurun_out/${tag}_class_w32_def.txt 2>&1; head -1 gpurun_out/${tag}_class_w32_def.txt
for v in "$@"; do
OQPB_LIB=openqp_b200/libopenqp_b200_$v.so timeout 600 python tools/class_profile.py w32 > gpurun_out/${tag}_class_w32_$v.txt 2>&1; head -1 gpurun_out/${tag}_class_w32_$v.txt
done
