"""Times the training-path kernels at BASELINE config 2 (B=8, T=2048, E=2560):
cg_rnn_scan_bwd (reads gy, a, h; writes dx, da: 5*s bytes per element) and
cg_conv1d_bwd (reads gy, x; writes dx: 3*s bytes per element) against the
measured HBM peak.  CUDA events around `iters` back-to-back calls, 3 repeats."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cadence_gemma_b200 import _abi  # noqa: E402


def timed(fn, iters=20):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  reps = []
  for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
      fn()
    b.record()
    b.synchronize()
    reps.append(a.elapsed_time(b) * 1e3 / iters)
  return sorted(reps)[1]


def main():
  peak = 6548.2
  try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
  except Exception:
    pass
  B, T, E = 8, 2048, 2560
  dev = "cuda"
  out = {"shape": [B, T, E], "hbm_peak_gbs": peak}
  for name, dtype in (("bf16", torch.bfloat16), ("f32", torch.float32)):
    s = 2 if dtype == torch.bfloat16 else 4
    gy = torch.randn(B, T, E, device=dev).to(dtype)
    a = (0.6 + 0.39 * torch.rand(B, T, E, device=dev)).to(dtype)
    h = torch.randn(B, T, E, device=dev).to(dtype)
    x = torch.randn(B, T, E, device=dev).to(dtype)
    w = torch.randn(4, E, device=dev).to(dtype)
    reset = torch.zeros(B, T, dtype=torch.bool, device=dev)
    reset[:, 0] = True
    seg = torch.arange(T, dtype=torch.int32, device=dev)[None].repeat(B, 1)
    h0 = torch.randn(B, E, device=dev)
    gl = torch.randn(B, E, device=dev)
    n = B * T * E
    us = timed(lambda: _abi.rnn_scan_bwd(gy, gl, a, h, reset, h0))
    out[f"rnn_scan_bwd_{name}"] = {"us": us, "GBps": 5 * s * n / us / 1e3, "frac": 5 * s * n / us / 1e3 / peak}
    for variant in ([0] if "--variants" not in sys.argv else range(6)):
      us = timed(lambda: _abi.rnn_scan_fwd(x, a, reset, h0, arith_mode=variant << 8))
      key = f"rnn_scan_fwd_{name}" + (f"_v{variant}" if variant else "")
      out[key] = {"us": us, "GBps": 3 * s * n / us / 1e3, "frac": 3 * s * n / us / 1e3 / peak}
    us = timed(lambda: _abi.conv1d_bwd(gy, x, w, seg))
    out[f"conv1d_bwd_{name}"] = {"us": us, "GBps": 3 * s * n / us / 1e3, "frac": 3 * s * n / us / 1e3 / peak}
  # one training step (forward + backward) of the RG-LRU module and of the whole
  # recurrent block at 2B shapes: training kernels vs the ATen op sequence around
  # the differentiable scan
  import cadence_gemma_b200 as cg
  from cadence_gemma_b200 import layers
  from cadence_gemma_b200.modules import RecurrentBlock
  dtype = torch.bfloat16
  lru = cg.RGLRU(E, 10, device=dev, dtype=dtype)
  blk = RecurrentBlock(width=2560, num_heads=10, lru_width=E, conv1d_temporal_width=4, device=dev, dtype=dtype)
  seg = torch.arange(T, dtype=torch.int32, device=dev)[None].repeat(B, 1)
  xin = torch.randn(B, T, E, device=dev).to(dtype).requires_grad_()
  gyo = torch.randn(B, T, E, device=dev).to(dtype)

  def train_step(mod):
    with torch.enable_grad():
      y, _ = mod(xin, seg)
      y.backward(gyo)

  for flag, tag in ((True, "kernels"), (False, "aten")):
    prev = layers.set_train_kernels(flag)
    out[f"rglru_train_step_{tag}_us"] = timed(lambda: train_step(lru), iters=5)
    out[f"recurrent_block_train_step_{tag}_us"] = timed(lambda: train_step(blk), iters=5)
    layers.set_train_kernels(prev)
  print(json.dumps(out))


if __name__ == "__main__":
  main()
