#!/bin/bash
out=gpurun_out/r3_ab6.log; : > $out
for uni in 0 1; do for fw in 0 1; do for sg in 0 300 1000 3000; do
echo "== uni $uni forward $fw stagger $sg" >> $out
CG_B200_UNI=$uni CG_B200_FORWARD=$fw CG_B200_STAGGER=$sg CG_B200_DYNAMIC=0 timeout 200 python - >> $out 2>&1 <<'P'
import sys, json, torch
sys.path.insert(0, ".")
import cadence_gemma_b200 as cg
from scripts import fused_check
E, H = 2560, 10
res = {}
for B, T in [(8, 2048), (16, 1024)]:
  x, lru, seg, _ = fused_check.make(B, T, E, H, resets=False)
  conv = cg.Conv1D(E, 4, device=x.device, dtype=torch.bfloat16)
  y = torch.empty_like(x); h = torch.empty((B, E), dtype=torch.float32, device=x.device); cs = torch.empty((B, 3, E), dtype=x.dtype, device=x.device)
  with torch.no_grad():
    conv.w.normal_(0, 0.4); conv.b.normal_(0, 0.2)
    run = lambda: cg.recurrent_hot_path(conv, lru, x, seg, out=y, last_h_out=h, conv_cache_out=cs)
    res[f"{B}x{T}"] = round(fused_check._time(run, 40), 1)
print(json.dumps(res))
P
done; done; done
grep -A1 "^==" $out | grep -v "^--" | paste - - | cut -c1-200
