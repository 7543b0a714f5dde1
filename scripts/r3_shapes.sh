#!/bin/bash
out=gpurun_out/r3_shapes.log; : > $out
timeout 600 python scripts/fused_check.py --case conv_shapes >> $out 2>&1
timeout 200 python - >> $out 2>&1 <<'P'
import sys
sys.path.insert(0, ".")
from scripts import fused_check
for B, T in [(16, 8192), (4, 8192), (16, 1024), (64, 512)]:
  fused_check.conv_case(B, T, iters=10, check=False)
P
cat $out
