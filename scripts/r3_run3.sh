#!/bin/bash
out=gpurun_out/r3_run3.log; : > $out
V=$PWD/cadence_gemma_b200/csrc/variants
echo "== trace, one launch" >> $out
CG_B200_LIB=$V/lib_trace.so timeout 120 python scripts/fused_trace.py --conv > /dev/null 2>>$out
python scripts/trace_stats.py >> $out 2>&1
P1="python scripts/profile_fused.py --B 8 --T 2048 --conv 1"
$P1 > gpurun_out/r3_plain_p1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rglru_fused -s 2 -c 1 -f -o gpurun_out/r3_prof_fused_conv $P1 > gpurun_out/r3_ncu_p1.log 2>&1
echo "fused+conv capture rc=$?" >> $out
cat $out
