#!/bin/bash
# usage: tools/ncu_kernel.sh <out-name> <kind: eri_kernel|eri_small_kernel> <la> <lb> <lc> <ld> <workload>
# one `ncu --set full` capture of one class kernel (B200_PROFILING.md recipe); output in gpurun_out/<out-name>.ncu-rep
out=$1; kind=$2; la=$3; lb=$4; lc=$5; ld=$6; wl=${7:-w8}
rx="${kind}<\\(int\\)${la}, \\(int\\)${lb}, \\(int\\)${lc}, \\(int\\)${ld}>"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:${rx}" -c 1 \
    -o gpurun_out/${out} -f python tools/run_build.py ${wl} 1 > gpurun_out/${out}.log 2>&1
tail -2 gpurun_out/${out}.log
