#!/bin/bash
out=gpurun_out/r2_run6.log; : > $out
V=$PWD/cadence_gemma_b200/csrc/variants
for rep in 1 2; do for v in base cst1 cst2; do
  echo "== $v conv" >> $out
  CG_B200_LIB=$V/lib_$v.so timeout 200 python scripts/fused_check.py --case conv 2>&1 | tail -1 | cut -c1-200 >> $out
done; done
echo "== bench config2" >> $out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "rc=$?" >> $out
python - >> $out <<'P'
import json
d=json.load(open("gpurun_out/r2_bench_n1.json"))
print({k:d[k] for k in ("value","ms_per_step","kernels_us","clocks")}, d["roofline_step"]["frac"], d["parity"]["shipped"]["us_per_step"])
P
cat $out
