#!/bin/bash
out=gpurun_out/r3_ab9.log; : > $out
V=$PWD/cadence_gemma_b200/csrc/variants
for rep in 1 2; do for v in "$@"; do
echo "== $v" >> $out
CG_B200_LIB=$V/lib_$v.so timeout 300 python - >> $out 2>&1 <<'P'
import sys, json, torch
sys.path.insert(0, ".")
import cadence_gemma_b200 as cg
from scripts import fused_check
E, H = 2560, 10
res = {}
for B, T in [(7, 2048), (8, 2048), (8, 1024), (8, 4096), (12, 2048), (16, 1024), (16, 2048), (6, 2048)]:
  x, lru, seg, _ = fused_check.make(B, T, E, H, resets=False)
  conv = cg.Conv1D(E, 4, device=x.device, dtype=torch.bfloat16)
  y = torch.empty_like(x); h = torch.empty((B, E), dtype=torch.float32, device=x.device); cs = torch.empty((B, 3, E), dtype=x.dtype, device=x.device)
  with torch.no_grad():
    conv.w.normal_(0, 0.4); conv.b.normal_(0, 0.2)
    run = lambda: cg.recurrent_hot_path(conv, lru, x, seg, out=y, last_h_out=h, conv_cache_out=cs)
    res[f"{B}x{T}"] = round(fused_check._time(run, 30), 1)
print(json.dumps(res))
P
done; done
grep -A1 "^==" $out | grep -v "^--" | paste - - | sort | cut -c1-300
