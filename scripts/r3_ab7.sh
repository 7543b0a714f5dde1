#!/bin/bash
out=gpurun_out/r3_ab7.log; : > $out
for dy in auto 0 auto; do
echo "== dynamic $dy" >> $out
if [ $dy = auto ]; then unset CG_B200_DYNAMIC; else export CG_B200_DYNAMIC=$dy; fi
timeout 300 python - >> $out 2>&1 <<'P'
import sys, json, torch
sys.path.insert(0, ".")
import cadence_gemma_b200 as cg
from scripts import fused_check
E, H = 2560, 10
res = {}
for B, T in [(8, 2048), (16, 1024), (6, 2048), (10, 2048), (12, 2048), (32, 768), (16, 2048), (8, 4096), (4, 8192), (16, 8192)]:
  x, lru, seg, _ = fused_check.make(B, T, E, H, resets=False)
  conv = cg.Conv1D(E, 4, device=x.device, dtype=torch.bfloat16)
  y = torch.empty_like(x); h = torch.empty((B, E), dtype=torch.float32, device=x.device); cs = torch.empty((B, 3, E), dtype=x.dtype, device=x.device)
  with torch.no_grad():
    conv.w.normal_(0, 0.4); conv.b.normal_(0, 0.2)
    run = lambda: cg.recurrent_hot_path(conv, lru, x, seg, out=y, last_h_out=h, conv_cache_out=cs)
    res[f"{B}x{T}"] = round(fused_check._time(run, 30 if B * T < 100000 else 10), 1)
print(json.dumps(res))
P
done
unset CG_B200_DYNAMIC
grep -A1 "^==" $out | grep -v "^--" | paste - - | cut -c1-300
