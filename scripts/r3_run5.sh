#!/bin/bash
out=gpurun_out/r3_run5.log; : > $out
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "fused" 2>&1 | tail -3 ) >> $out
timeout 300 python - >> $out 2>&1 <<'P'
import sys
sys.path.insert(0, ".")
from scripts import fused_check
for B, T in [(8, 2048), (8, 2048), (2, 8192), (4, 8192), (32, 768), (16, 8192), (1, 2048)]:
  fused_check.conv_case(B, T, iters=20, check=False)
P
cat $out
