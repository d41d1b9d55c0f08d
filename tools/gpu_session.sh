#!/bin/bash
# GPU box session: tests, bench, class profile and targeted ncu captures.  usage: tools/gpu_session.sh <tag> [steps...]
# Everything lands in gpurun_out/<tag>_*.  Steps: tests bench prof ncu (default: all)
tag=$1; shift
steps=${@:-tests bench prof ncu}
mkdir -p gpurun_out
for s in $steps; do
  case $s in
    tests) timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log ;;
    bench) timeout 900 python bench.py > gpurun_out/${tag}_bench_w32.json 2> gpurun_out/${tag}_bench_w32.err; cat gpurun_out/${tag}_bench_w32.json | head -c 600; echo ;;
    prof) timeout 600 python tools/class_profile.py w32 > gpurun_out/${tag}_class_w32.txt 2>&1; head -3 gpurun_out/${tag}_class_w32.txt ;;
    ncu)
      cap() {  # cap <name> <bra list> <ket list> <kernel regex>
        OQPB_ONLY="$2,$3" timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
          -k "regex:$4" -c 1 -o gpurun_out/${tag}_$1 -f python tools/run_build.py w32 1 > gpurun_out/${tag}_$1.log 2>&1
        tail -1 gpurun_out/${tag}_$1.log
      }
      cap small_2010_b00 12 4 'eri_small_kernel<\(int\)2, \(int\)0, \(int\)1, \(int\)0,'
      cap small_1000_b00 4 0 'eri_small_kernel<\(int\)1, \(int\)0, \(int\)0, \(int\)0,'
      cap small_1000_b12 5 2 'eri_small_kernel<\(int\)1, \(int\)0, \(int\)0, \(int\)0,'
      cap med_2120_b00 16 12 'eri_small_kernel<\(int\)2, \(int\)1, \(int\)2, \(int\)0,'
      cap grp_2111_b00 16 8 'eri_group_kernel<\(int\)2, \(int\)1, \(int\)1, \(int\)1,'
      cap grp_2221_b00 20 16 'eri_group_kernel<\(int\)2, \(int\)2, \(int\)2, \(int\)1,'
      ;;
  esac
done
