#!/bin/bash
out=gpurun_out/r3_trace2.log; : > $out
V=$PWD/cadence_gemma_b200/csrc/variants
for st in 0 1; do
echo "== one launch, steal $st" >> $out
CG_B200_STEAL=$st CG_B200_LIB=$V/lib_trace.so timeout 120 python scripts/fused_trace.py --conv 2>>$out | grep "kernel span\|^unit" >> $out
done
cat $out
