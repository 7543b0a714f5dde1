"""Minimal launcher for ncu: runs one kernel configuration a few times.

    python scripts/profile_target.py --kernel rglru --mode 2 --variant 1 [--iters 3]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cadence_gemma_b200 import _abi  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--kernel", default="rglru", choices=["rglru", "conv1d", "rnn_scan"])
  ap.add_argument("--mode", type=int, default=2)
  ap.add_argument("--variant", type=int, default=0)
  ap.add_argument("--dtype", default="bf16")
  ap.add_argument("--B", type=int, default=8)
  ap.add_argument("--T", type=int, default=2048)
  ap.add_argument("--E", type=int, default=2560)
  ap.add_argument("--iters", type=int, default=3)
  args = ap.parse_args()
  dev = "cuda:0"
  dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
  B, T, E = args.B, args.T, args.E
  g = torch.Generator(device=dev).manual_seed(1)
  x = torch.randn((B, T, E), device=dev, generator=g).to(dtype)
  gx = (torch.randn((B, T, E), device=dev, generator=g) * 1.5).to(dtype)
  ga = (torch.randn((B, T, E), device=dev, generator=g) * 1.5).to(dtype)
  ap_ = (torch.rand((E,), device=dev, generator=g) * -7 + 1).to(dtype)
  bx = torch.randn((E,), device=dev, generator=g).to(dtype)
  ba = torch.randn((E,), device=dev, generator=g).to(dtype)
  w = (torch.randn((4, E), device=dev, generator=g) * 0.5).to(dtype)
  seg = torch.arange(T, dtype=torch.int32, device=dev)[None].repeat(B, 1)
  rs = torch.zeros((B, T), dtype=torch.bool, device=dev)
  a = torch.rand_like(x)
  y = torch.empty_like(x)
  am = args.mode | (args.variant << 8)
  for _ in range(args.iters):
    if args.kernel == "rglru":
      _abi.rglru_fwd(x, gx, ga, bx, ba, ap_, seg, arith_mode=am, out=y)
    elif args.kernel == "conv1d":
      _abi.conv1d_fwd(x, w, bx, seg, arith_mode=am)
    else:
      _abi.rnn_scan_fwd(x, a, rs, None, arith_mode=am)
  torch.cuda.synchronize()
  print("ok")


if __name__ == "__main__":
  main()
