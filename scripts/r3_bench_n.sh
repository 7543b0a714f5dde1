#!/bin/bash
# bench.py at N GPUs with the final library (config 2, weak scaling)
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r3_bench_n$N.json 2> gpurun_out/r3_bench_n$N.err; echo "rc=$?"
python - $N <<'P'
import json,sys
n=sys.argv[1]
d=json.loads(open(f"gpurun_out/r3_bench_n{n}.json").read().strip().splitlines()[-1])
print(n, round(d["value"]/1e6,1), d["ms_per_step"], (d.get("with_gather_joined") or {}).get("ms_per_step"), d["clocks"].get("sm_mhz"), d["clocks"].get("samples"), round(d["e2e"]["value"]/1e6,2), round(d["e2e"]["ceiling"]["value"]/1e6,2))
P
