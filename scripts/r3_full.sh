#!/bin/bash
# full GPU suite + smoke + bench on the current tree
out=gpurun_out/r3_full.log; : > $out
( timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6 ) >> $out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 ) >> $out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3_bench_n1.json 2> gpurun_out/r3_bench_n1.err; echo "bench rc=$?" >> $out
python - >> $out <<'P'
import json
d=json.loads(open("gpurun_out/r3_bench_n1.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","kernels_us","route","gpu_launches")}, d["clocks"].get("sm_mhz"), d["clocks"].get("samples"), d["roofline_step"]["frac"], d["roofline"]["frac"])
print(json.dumps(d["e2e"])[:400])
print(json.dumps(d.get("parity"), indent=0)[:1200])
P
cat $out
