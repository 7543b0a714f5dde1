"""Launches the fused decode step and the three-launch path a few times (for ncu)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cadence_gemma_b200 as cg

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
E, H = 2560, 10
dev = "cuda"
torch.manual_seed(0)
conv = cg.Conv1D(E, 4, device=dev, dtype=torch.bfloat16)
lru = cg.RGLRU(E, H, device=dev, dtype=torch.bfloat16)
x = torch.randn(B, 1, E, device=dev).to(torch.bfloat16)
cache = torch.randn(B, 3, E, device=dev).to(torch.bfloat16)
h0 = torch.randn(B, E, device=dev)
seg = torch.full((B, 1), 16, device=dev, dtype=torch.int32)
with torch.no_grad():
  for _ in range(5):
    cg.recurrent_hot_path(conv, lru, x, seg, conv_cache=cache, lru_cache=h0)
  cg.set_fused(False)
  for _ in range(5):
    cg.recurrent_hot_path(conv, lru, x, seg, conv_cache=cache, lru_cache=h0)
torch.cuda.synchronize()
print("ok")
