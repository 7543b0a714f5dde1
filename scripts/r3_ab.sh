#!/bin/bash
# A/B of library variants on the one-launch route (config 2): scripts/r3_ab.sh name1 name2 ...
out=gpurun_out/r3_ab.log; : > $out
V=$PWD/cadence_gemma_b200/csrc/variants
for rep in 1 2; do
  for v in "$@"; do
    echo "== $v" >> $out
    CG_B200_LIB=$V/lib_$v.so timeout 120 python - >> $out 2>&1 <<'P'
import sys
sys.path.insert(0, ".")
from scripts import fused_check
fused_check.conv_case(check=False)
P
  done
done
grep -A1 "^==" $out | grep -v "^--" | paste - - | cut -c1-200
