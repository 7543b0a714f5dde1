#!/bin/bash
# round 2, 2-GPU run: bench config2 / config3 at N=2 + BatchShardedHotPath over NCCL
N=${1:-2}
out=gpurun_out/r3_run5_n$N.log; : > $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== multi_gpu_check" >> $out
timeout 600 $TR --master-port 29533 scripts/multi_gpu_check.py > gpurun_out/r3_multi_gpu_n$N.json 2>> $out; echo "rc=$?" >> $out
echo "== bench config2 N=$N" >> $out
timeout 600 $TR --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r3_bench_n$N.json 2> gpurun_out/r3_bench_n$N.err; echo "rc=$?" >> $out
echo "== bench config3 N=$N" >> $out
timeout 900 $TR --master-port 29535 bench.py --gpus $N --workload config3 --steps 5 --warmup 3 > gpurun_out/r3_config3_n$N.json 2> gpurun_out/r3_config3_n$N.err; echo "rc=$?" >> $out
tail -c 1500 $out
cat gpurun_out/r3_multi_gpu_n$N.json
python - $N <<'P'
import json,sys
n=sys.argv[1]
for f in (f"gpurun_out/r3_bench_n{n}.json", f"gpurun_out/r3_config3_n{n}.json"):
  try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, {k:d[k] for k in ("value","ms_per_step","n_gpus") if k in d}, json.dumps(d.get("with_gather_joined"))[:200], json.dumps(d.get("e2e"))[:160], json.dumps(d.get("recurrent_hot_path"))[:200])
  except Exception as e:
    print(f, "ERR", e, open(f.replace(".json",".err")).read()[-1200:])
P
