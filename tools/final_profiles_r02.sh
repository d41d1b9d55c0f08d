#!/bin/bash
# Round-2 measurement session on ONE B200: tests, bench lines of every configuration, launch list, HBM traffic, ncu captures
# of the three kernel families, per-class profile, SCF-density (dD) regime, sigma session.  Everything under gpurun_out/r02_*.
mkdir -p gpurun_out
P=gpurun_out
nvidia-smi -L
timeout 1700 python -m pytest tests -m gpu -x -q -s > $P/r02_tests.log 2>&1; tail -3 $P/r02_tests.log
b() { name=$1; shift; timeout 900 python bench.py "$@" > $P/r02_bench_$name.json 2> $P/r02_bench_$name.err; python -c "
import json; d=json.load(open('$P/r02_bench_$name.json')); print('$name', round(d['ms_per_step'],2), 'ms', d.get('roofline',{}).get('frac'), d.get('cpu_baseline',{}).get('value'))"; }
b w32 --steps 5 --warmup 3
b w32_reference --impl reference --steps 1 --warmup 1
b w32_cam --cam --steps 3 --warmup 3 --no-cpu-baseline
b c1 --workload c1 --steps 20 --warmup 5
b c2 --workload c2 --steps 20 --warmup 5
b c3 --workload c3 --steps 10 --warmup 3
b c5 --workload c5 --steps 3 --warmup 3
b c5_cut1e-8 --workload c5 --cutoff 1e-8 --steps 3 --warmup 3 --no-cpu-baseline
b c4 --workload c4 --steps 2 --warmup 3 --no-cpu-baseline
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $P/r02_launches_w32.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $P/r02_launches_w32.log 2>&1
python tools/launch_summary.py $P/r02_launches_w32.csv "python bench.py --steps 1 --warmup 1 --no-cpu-baseline" > $P/r02_launches_w32_summary.txt; gzip -f $P/r02_launches_w32.csv; head -12 $P/r02_launches_w32_summary.txt
timeout 900 python tools/dram_traffic.py w32 > $P/r02_dram_w32.log 2>&1; tail -2 $P/r02_dram_w32.log; cp profiles/r02_dram_w32.json $P/ 2>/dev/null
python tools/ncu_top.py w32 r02_small_1000 "eri_run_kernel<\(int\)1, \(int\)0, \(int\)0, \(int\)0," \
    r02_med_2120 "eri_run_kernel<\(int\)2, \(int\)1, \(int\)2, \(int\)0," r02_kown_2111 "eri_kown_kernel<\(int\)2, \(int\)1, \(int\)1, \(int\)1," 2>&1 | tail -9
timeout 600 python tools/class_profile.py w32 > $P/r02_class_profile_w32.txt 2>&1; head -1 $P/r02_class_profile_w32.txt
timeout 600 python tools/sigma_bench.py c5 12 1e-8 > $P/r02_sigma_c5.txt 2>&1; cat $P/r02_sigma_c5.txt
timeout 600 python tools/sigma_bench.py c5 12 5e-11 >> $P/r02_sigma_c5.txt 2>&1
timeout 1200 python tools/scf_density_bench.py w32 $P/r02_scf_w32.json > $P/r02_scf_w32.txt 2>&1; tail -4 $P/r02_scf_w32.txt
