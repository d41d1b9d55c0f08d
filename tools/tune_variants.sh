#!/bin/bash
# Build tuning variants of the library (register caps per kernel family) next to the default one:
#   tools/tune_variants.sh          -> libopenqp_b200_v1.so, libopenqp_b200_v2.so
# then on the GPU box: OQPB_LIB=openqp_b200/libopenqp_b200_v1.so python tools/class_profile.py w32
set -e
cd "$(dirname "$0")/.."
OQPB_VARIANT=v1 OQPB_EXTRA_FLAGS="-DOQPB_SMALL_REGS=128 -DOQPB_MED_REGS=168 -DOQPB_GRP_REGS=168" python -m openqp_b200.build
OQPB_VARIANT=v2 OQPB_EXTRA_FLAGS="-DOQPB_SMALL_REGS=96 -DOQPB_MED_REGS=128 -DOQPB_GRP_REGS=128" python -m openqp_b200.build
