#!/bin/bash
# GPU box: launch list of one w32 Fock build + `--set full` captures of the top kernels -> gpurun_out/
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_w32.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01_launches_w32.log 2>&1
python tools/ncu_top.py w32 r01f_small_1000 "eri_small_kernel<\(int\)1, \(int\)0, \(int\)0, \(int\)0," \
    r01f_grp_2111 "eri_group_kernel<\(int\)2, \(int\)1, \(int\)1, \(int\)1," r01f_med_2120 "eri_small_kernel<\(int\)2, \(int\)1, \(int\)2, \(int\)0," 2>&1 | tail -6
OQPB_NCU_TARGET=tools/mrsf_bench.py OQPB_NCU_ARGS="12 1" python tools/ncu_top.py c5 r01f_mrsf_grp_2111 \
    "eri_group_kernel<\(int\)2, \(int\)1, \(int\)1, \(int\)1," 2>&1 | tail -3
