#!/bin/bash
out=gpurun_out/r3_trace4.log; : > $out
V=$PWD/cadence_gemma_b200/csrc/variants
for v in trace0 trace1; do
echo "== $v" >> $out
CG_B200_LIB=$V/lib_$v.so timeout 120 python scripts/fused_trace.py --conv 2>>$out | grep "kernel span" >> $out
python scripts/trace_stats.py >> $out 2>&1
done
cat $out
