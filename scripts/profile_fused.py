"""Minimal launcher for ncu: the hot path of one shape a few times.

    python scripts/profile_fused.py --B 8 --T 2048 --conv 0     (Conv1D kernel + fused tcgen05 kernel)
    python scripts/profile_fused.py --B 2 --T 2048 --conv 1     (ONE launch: convolution inside the fused kernel)
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cadence_gemma_b200 as cg  # noqa: E402
from cadence_gemma_b200 import pipeline  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--B", type=int, default=8)
  ap.add_argument("--T", type=int, default=2048)
  ap.add_argument("--conv", type=int, default=0)
  ap.add_argument("--iters", type=int, default=4)
  a = ap.parse_args()
  dev = torch.device("cuda", 0)
  E, H = 2560, 10
  torch.manual_seed(1)
  conv = cg.Conv1D(E, 4, device=dev, dtype=torch.bfloat16)
  lru = cg.RGLRU(E, H, device=dev, dtype=torch.bfloat16)
  with torch.no_grad():
    conv.w.normal_(0, 0.4); conv.b.normal_(0, 0.2)
    lru.input_gate.b.normal_(); lru.a_gate.b.normal_()
    x = torch.randn(a.B, a.T, E, device=dev).to(torch.bfloat16)
    seg = torch.arange(a.T, device=dev, dtype=torch.int32)[None].repeat(a.B, 1)
    pipeline.set_fused_conv(bool(a.conv))
    for _ in range(a.iters):
      cg.recurrent_hot_path(conv, lru, x, seg)
  torch.cuda.synchronize()
  print("ok")


if __name__ == "__main__":
  main()
