#!/bin/bash
out=gpurun_out/r2_run4.log; : > $out
V=$PWD/cadence_gemma_b200/csrc/variants
for rep in 1 2; do for v in base olddesc; do
  echo "== $v time" >> $out
  CG_B200_LIB=$V/lib_$v.so timeout 200 python scripts/fused_check.py --case time 2>&1 | tail -1 | cut -c1-120 >> $out
done; done
echo "== cfg2 check (base)" >> $out
CG_B200_LIB=$V/lib_base.so timeout 200 python scripts/fused_check.py --case cfg2 2>&1 | tail -1 | cut -c1-400 >> $out
echo "== conv shapes" >> $out
timeout 300 python scripts/fused_check.py --case conv_shapes >> $out 2>&1
echo "== decode test" >> $out
( timeout 600 python -m pytest tests/test_gpu_decode.py tests/test_gpu_parity.py -x -q -k "decode or fused_rglru_golden or fused_conv" 2>&1 | tail -5 ) >> $out
echo "== bench_decode" >> $out
timeout 600 python scripts/bench_decode.py > gpurun_out/r2_decode.json 2>> $out; cat gpurun_out/r2_decode.json >> $out
echo "== bench config2" >> $out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "rc=$?" >> $out
python - >> $out <<'P'
import json
d=json.load(open("gpurun_out/r2_bench_n1.json"))
print({k:d[k] for k in ("value","ms_per_step","kernels_us","clocks")}, d["roofline_step"]["frac"], d["parity"]["shipped"]["us_per_step"])
P
cat $out
