#!/bin/bash
# round 2, GPU run 1: first correctness + timing of the in-kernel convolution
out=gpurun_out/r2_run1.log; : > $out
V=$PWD/cadence_gemma_b200/csrc/variants
run() { echo "== $*" >> $out; timeout 300 "$@" >> $out 2>&1; echo "rc=$?" >> $out; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $out
run python scripts/fused_check.py --case small
run python scripts/fused_check.py --case ragged
run python scripts/fused_check.py --case conv
for v in base nomax; do
  for rep in 1 2; do
    echo "== variant $v time" >> $out
    CG_B200_LIB=$V/lib_$v.so timeout 120 python scripts/fused_check.py --case time 2>&1 | tail -1 >> $out
    echo "== variant $v conv" >> $out
    CG_B200_LIB=$V/lib_$v.so timeout 120 python scripts/fused_check.py --case conv 2>&1 | tail -2 >> $out
  done
done
run python -m pytest tests/test_gpu_parity.py -x -q -k "fused_conv or fused_rglru_golden or fused_rglru_equals or family_loop or recurrent_block_golden"
tail -60 $out
