#!/bin/bash
# A/B of build variants (F fused partners, S stages; scripts/build_variants.py) and threads per
# chain on config 4.  Usage: variant_ab.sh "lib:tpc" ...   (lib "" = the in-tree default build)
V=nutpie_b200/variants
run() { echo "=== lib=${1:-default} tpc=$2"; NB200_LIB=$1 TPC=$2 MODES=${MODES:-3} REPS=${REPS:-1} timeout 120 python scripts/stage_ab.py 2>&1 | grep -E "kernel_ms|^mode|rror" | cut -c1-230; }
for spec in "$@"; do
  lib=${spec%%:*}; tpc=${spec##*:}
  [ -n "$lib" ] && lib=$V/libnutpie_b200_$lib.so
  run "$lib" $tpc
done
