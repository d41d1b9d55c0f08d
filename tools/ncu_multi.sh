#!/bin/bash
# usage: tools/ncu_multi.sh <workload> <count> <out-name> <kernel-regex>
# `ncu --set full` capture of up to <count> launches whose demangled name matches <kernel-regex>
wl=$1; cnt=$2; out=$3; rx=$4
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:${rx}" -c ${cnt} \
    -o gpurun_out/${out} -f python tools/run_build.py ${wl} 1 > gpurun_out/${out}.log 2>&1
tail -2 gpurun_out/${out}.log
