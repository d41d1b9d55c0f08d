#!/bin/bash
# ket-owner group kernel: parity tests with it on every class it covers, then per-class A/B on w32
tag=${1:-s6}
mkdir -p gpurun_out
nvidia-smi -L
OQPB_KOWN=2 timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
OQPB_KOWN=0 timeout 600 python tools/class_profile.py w32 > gpurun_out/${tag}_class_w32_kown0.txt 2>&1; head -1 gpurun_out/${tag}_class_w32_kown0.txt
OQPB_KOWN=2 timeout 600 python tools/class_profile.py w32 > gpurun_out/${tag}_class_w32_kown2.txt 2>&1; head -1 gpurun_out/${tag}_class_w32_kown2.txt
