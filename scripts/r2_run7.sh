#!/bin/bash
out=gpurun_out/r2_run7.log; : > $out
( timeout 900 python -m pytest tests/test_gpu_configs.py -x -q -s -k "packed or config2" 2>&1 | tail -25 ) >> $out
echo "== time mode 2 (FAST) / 8 (PACKED) / 0 (exact)" >> $out
for m in 2 8 0; do timeout 200 python scripts/fused_check.py --case time --mode $m 2>&1 | tail -1 | cut -c1-130 >> $out; done
echo "== conv case PACKED (one launch vs two kernels)" >> $out
CG_FC_MODE=8 timeout 200 python - >> $out 2>&1 <<'P'
import sys, torch
sys.path.insert(0, ".")
import cadence_gemma_b200 as cg
from scripts import fused_check
cg.set_arith_mode(8)
fused_check.conv_case(check=False)
fused_check.conv_case(2, 2048, check=False)
P
echo "== bench" >> $out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "rc=$?" >> $out
python - >> $out <<'P'
import json
d=json.load(open("gpurun_out/r2_bench_n1.json"))
print({k:d[k] for k in ("value","ms_per_step","kernels_us")}, d["clocks"]["sm_mhz"], d["clocks"]["ms_per_step_of_the_sampled_replica"], d["roofline_step"]["frac"])
print(json.dumps(d["parity"], indent=0)[:1500])
P
cat $out
