#!/bin/bash
# tile descriptors: correctness of both fused routes, then A/B timing
out=gpurun_out/r3_run4.log; : > $out
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "fused" 2>&1 | tail -4 ) >> $out
( timeout 600 python -m pytest tests/test_gpu_configs.py -x -q 2>&1 | tail -3 ) >> $out
for i in 1 2; do timeout 200 python scripts/fused_check.py --case time 2>&1 | tail -1 | cut -c1-120 >> $out; done
timeout 200 python scripts/fused_check.py --case conv >> $out 2>&1
echo "rc=$?" >> $out
cat $out
