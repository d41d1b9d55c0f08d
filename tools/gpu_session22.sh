#!/bin/bash
tag=${1:-s22}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -2 gpurun_out/${tag}_tests.log
timeout 600 python tools/class_profile.py w32 > gpurun_out/${tag}_class_w32_def.txt 2>&1; head -1 gpurun_out/${tag}_class_w32_def.txt
OQPB_KOWN=2 timeout 600 python tools/class_profile.py w32 > gpurun_out/${tag}_class_w32_k2.txt 2>&1; head -1 gpurun_out/${tag}_class_w32_k2.txt
OQPB_KOWN=2 OQPB_LIB=openqp_b200/libopenqp_b200_r168.so timeout 600 python tools/class_profile.py w32 > gpurun_out/${tag}_class_w32_k2r168.txt 2>&1; head -1 gpurun_out/${tag}_class_w32_k2r168.txt
