#!/bin/bash
tag=${1:-s7}
mkdir -p gpurun_out
OQPB_KOWN=2 timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
rx='eri_kown_kernel<\(int\)2, \(int\)1, \(int\)1, \(int\)1'
OQPB_KOWN=2 OQPB_ONLY=16,8 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:${rx}" -c 1 \
    -o gpurun_out/${tag}_kown_2111 -f python tools/run_build.py w32 1 > gpurun_out/${tag}_kown_2111.log 2>&1
tail -2 gpurun_out/${tag}_kown_2111.log
rx='eri_group_kernel<\(int\)2, \(int\)1, \(int\)1, \(int\)1'
OQPB_KOWN=0 OQPB_ONLY=16,8 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:${rx}" -c 1 \
    -o gpurun_out/${tag}_grp_2111 -f python tools/run_build.py w32 1 > gpurun_out/${tag}_grp_2111.log 2>&1
tail -2 gpurun_out/${tag}_grp_2111.log
