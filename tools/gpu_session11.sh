#!/bin/bash
tag=${1:-s11}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s -k "sigma" > gpurun_out/${tag}_tests_sigma.log 2>&1; tail -5 gpurun_out/${tag}_tests_sigma.log
timeout 600 python tools/class_profile.py w32 > gpurun_out/${tag}_class_w32.txt 2>&1; head -1 gpurun_out/${tag}_class_w32.txt
OQPB_LIB=openqp_b200/libopenqp_b200_nosegc.so timeout 600 python tools/class_profile.py w32 > gpurun_out/${tag}_class_w32_nosegc.txt 2>&1; head -1 gpurun_out/${tag}_class_w32_nosegc.txt
