#!/bin/bash
out=gpurun_out/r3_trace3.log; : > $out
V=$PWD/cadence_gemma_b200/csrc/variants
CG_B200_LIB=$V/lib_trace.so timeout 120 python scripts/fused_trace.py --conv 2>>$out | grep "kernel span\|^unit" >> $out
python scripts/trace_stats.py >> $out 2>&1
cat $out
