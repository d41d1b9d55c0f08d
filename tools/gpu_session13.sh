#!/bin/bash
tag=${1:-s13}
mkdir -p gpurun_out
OQPB_LIB=openqp_b200/libopenqp_b200_mvol.so timeout 600 python tools/class_profile.py w32 > gpurun_out/${tag}_class_w32_mvol.txt 2>&1; head -1 gpurun_out/${tag}_class_w32_mvol.txt
timeout 900 python tools/partition_bench.py w32 2,4,8 > gpurun_out/${tag}_partition_w32.txt 2>&1; cat gpurun_out/${tag}_partition_w32.txt
OQPB_NLANES=8 timeout 900 python tools/partition_bench.py w32 8 > gpurun_out/${tag}_partition_w32_l8.txt 2>&1; cat gpurun_out/${tag}_partition_w32_l8.txt
timeout 900 python tools/scf_density_bench.py c3 gpurun_out/${tag}_scf_c3.json > gpurun_out/${tag}_scf_c3.txt 2>&1; tail -5 gpurun_out/${tag}_scf_c3.txt
