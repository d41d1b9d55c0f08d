#!/bin/bash
# Build tuning variants of the library next to the default one:  tools/tune_variants.sh name "flags" [name "flags" ...]
#   e.g. tools/tune_variants.sh v1 "-DOQPB_GRP_LIMIT=56" v2 "-DOQPB_SMALL_REGS=128"
# then on the GPU box: OQPB_LIB=openqp_b200/libopenqp_b200_v1.so python tools/class_profile.py w32
set -e
cd "$(dirname "$0")/.."
while [ $# -ge 2 ]; do
  OQPB_VARIANT=$1 OQPB_EXTRA_FLAGS="$2" python -m openqp_b200.build
  shift 2
done
