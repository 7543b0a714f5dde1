#!/bin/bash
# first GPU run of the epilogue-warpgroup convolution (2-CTA cluster exchange)
out=gpurun_out/r3_run1.log; : > $out
echo "== small shapes: one launch vs two kernels (identical bits expected)" >> $out
timeout 120 python - >> $out 2>&1 <<'P'
import sys, torch, json
sys.path.insert(0, ".")
from scripts import fused_check
for B, T in [(1, 64), (2, 128), (1, 2048), (8, 512)]:
  fused_check.conv_case(B, T, iters=5, check=True)
P
echo "rc=$?" >> $out
echo "== conv case config 2" >> $out
timeout 200 python scripts/fused_check.py --case conv >> $out 2>&1
echo "rc=$?" >> $out
( timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fused_conv_prefill" 2>&1 | tail -8 ) >> $out
cat $out
