#!/bin/bash
tag=${1:-s9}
mkdir -p gpurun_out
rx='eri_kown_kernel<\(int\)2, \(int\)1, \(int\)1, \(int\)1'
OQPB_LIB=openqp_b200/libopenqp_b200_a20.so OQPB_KOWN=2 OQPB_ONLY=16,8 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:${rx}" -c 1 \
    -o gpurun_out/${tag}_kown_2111 -f python tools/run_build.py w32 1 > gpurun_out/${tag}_kown_2111.log 2>&1
tail -2 gpurun_out/${tag}_kown_2111.log
