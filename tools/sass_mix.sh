#!/bin/bash
# sass_mix.sh <object-or-so> <mangled-kernel-name-regex>: SASS instruction mix of the matching kernels
# (counts of the opcodes that decide the FP64 / shared-memory / local-memory balance of the ERI kernels)
f=$1; pat=$2
cuobjdump -sass "$f" 2>/dev/null | awk -v pat="$pat" '
/Function :/ { name=$3; on = (name ~ pat); if (on) { n++; names[n]=name } ; next }
on && /^ +\/\*[0-9a-f]+\*\/ +[A-Z@]/ {
  op=$2; if (op ~ /^@/) op=$3; sub(/\..*/, "", op); sub(/;$/, "", op); cnt[names[n] SUBSEP op]++; tot[names[n]]++ }
END { for (i=1;i<=n;i++) { nm=names[i]; printf "%s\n  total %d", nm, tot[nm];
  split("DFMA DMUL DADD DMMA LDS STS LDL STL LDG RED ATOMG SHFL IMAD IADD3 MOV ISETP BRA BAR MUFU", ops, " ");
  for (j=1;j<=19;j++) if (cnt[nm SUBSEP ops[j]]) printf "  %s %d", ops[j], cnt[nm SUBSEP ops[j]]; printf "\n" } }' | c++filt
