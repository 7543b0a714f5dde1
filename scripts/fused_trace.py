"""Timeline of one CTA of the fused kernel (needs a -DCGF_TRACE=1 build, see
scripts/build_variants.py): per role, the clock64 of every pipeline event."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cadence_gemma_b200 import _abi  # noqa: E402
from scripts.fused_check import make  # noqa: E402

B, T, E, H = 8, 2048, 2560, 10
x, lru, seg, _ = make(B, T, E, H, resets=False)
wpack = _abi.pack_gate_weights(lru.input_gate.w, lru.a_gate.w)
ws = _abi.fused_workspace(x.device, B, T, E)
conv = "--conv" in sys.argv          # the one-launch route (convolution inside the kernel)
cw = (torch.randn(4, E, device=x.device) * 0.4).to(torch.bfloat16)
cb = (torch.randn(E, device=x.device) * 0.2).to(torch.bfloat16)
for _ in range(3):
  if conv:
    out = _abi.recurrent_prefill_fwd(x, cw, cb, wpack, lru.input_gate.b, lru.a_gate.b, lru.a_param, seg, H,
                                     arith_mode=2, debug=True, workspace=ws)
    out = (out[0], out[2], out[3])
  else:
    out = _abi.rglru_fused_fwd(x, wpack, lru.input_gate.b, lru.a_gate.b, lru.a_param, seg, H,
                               arith_mode=2, debug=True, workspace=ws)
torch.cuda.synchronize()
tr = out[2].view(-1).view(torch.int64)[: 7 * 1024].view(7, 1024).cpu()
names = ["producer", "mma", "wg0", "wg1", "wg2", "wg3", "conv"]
t0 = None
res = {}
for r in range(7):
  ev = [(int(v) >> 4, int(v) & 15) for v in tr[r].tolist() if v != 0]
  if ev and (t0 is None or ev[0][0] < t0):
    t0 = ev[0][0]
  res[names[r]] = ev
for k, ev in res.items():
  res[k] = [(t - t0, c) for t, c in ev]
json.dump(res, open("gpurun_out/fused_trace.json", "w"))
for k, ev in res.items():
  print(k, len(ev), ev[:40])

# per-CTA start / end times (ns): which CTAs set the kernel's time?
tt = out[2].view(-1).view(torch.int64)[7 * 1024: 7 * 1024 + 512].cpu()
end, start = tt[:148].tolist(), tt[256:256 + 148].tolist()
t0g = min(start)
dur = [(e - t0g) / 1e3 for e in end]
fam = lambda i: (i // 2) % 10 if conv else i % 20
import collections
by = collections.defaultdict(list)
for i, d in enumerate(dur):
  by[fam(i)].append(d)
print("kernel span us", round(max(dur), 1), "start skew us", round((max(start) - t0g) / 1e3, 1))
for f in sorted(by):
  print("unit", f, "ctas", len(by[f]), "end us min/max", round(min(by[f]), 1), round(max(by[f]), 1))
