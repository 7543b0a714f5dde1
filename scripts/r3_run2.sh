#!/bin/bash
# barrier-free epilogue convolution: correctness, A/B timing, one full ncu capture of the one-launch kernel
out=gpurun_out/r3_run2.log; : > $out
( timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fused_conv_prefill" 2>&1 | tail -4 ) >> $out
timeout 200 python scripts/fused_check.py --case conv >> $out 2>&1
echo "rc=$?" >> $out
P1="python scripts/profile_fused.py --B 8 --T 2048 --conv 1"
$P1 > gpurun_out/r3_plain_p1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rglru_fused -s 2 -c 1 -f -o gpurun_out/r3_prof_fused_conv $P1 > gpurun_out/r3_ncu_p1.log 2>&1
echo "fused+conv capture rc=$?" >> $out
cat $out
