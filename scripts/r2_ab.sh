#!/bin/bash
# A/B timing of library variants (scripts/build_variants.py) on a GPU box:
#   scripts/r2_ab.sh <case> name1 name2 ...   -> gpurun_out/r2_ab.log
case=$1; shift
out=gpurun_out/r2_ab.log; : > $out
V=$PWD/cadence_gemma_b200/csrc/variants
for rep in 1 2; do
  for v in "$@"; do
    echo "== $v $case" >> $out
    CG_B200_LIB=$V/lib_$v.so timeout 200 python scripts/fused_check.py --case $case 2>&1 | tail -3 >> $out
  done
done
cat $out
