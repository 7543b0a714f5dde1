#!/bin/bash
# round 2, GPU run 3: GPU suite + bench config2 + bench config3 (model level)
out=gpurun_out/r2_run3.log; : > $out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) >> $out
echo "== bench config2" >> $out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "rc=$?" >> $out
echo "== bench config3" >> $out
timeout 900 python bench.py --workload config3 --steps 5 --warmup 3 > gpurun_out/r2_config3_n1.json 2> gpurun_out/r2_config3_n1.err; echo "rc=$?" >> $out
tail -c 3000 $out
python - <<'P'
import json
for f in ("gpurun_out/r2_bench_n1.json","gpurun_out/r2_config3_n1.json"):
  try:
    d=json.load(open(f))
    print(f, {k:d[k] for k in ("value","ms_per_step","gpu_launches") if k in d}, json.dumps(d.get("kernels_us")), json.dumps(d.get("recurrent_hot_path")), json.dumps(d.get("e2e"))[:300])
  except Exception as e:
    print(f, "ERR", e, open(f.replace(".json",".err")).read()[-1500:])
P
