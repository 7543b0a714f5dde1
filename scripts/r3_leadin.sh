#!/bin/bash
N=${1:-2}
out=gpurun_out/r3_leadin_n$N.log; : > $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for li in 4 16 4 16; do
  CG_BENCH_LEAD_IN=$li timeout 600 $TR --master-port 295$((30 + li)) bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('lead_in $li', d['n_gpus'], d['value'], d['ms_per_step'], (d.get('with_gather_joined') or {}).get('ms_per_step'), d['clocks'].get('sm_mhz'))" >> $out
done
cat $out
