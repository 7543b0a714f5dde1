"""Development check of the fused tcgen05 RG-LRU kernel against the unfused
path (cuBLAS gate GEMM + scan kernel) on a GPU box.

    python scripts/fused_check.py --case small|ragged|cfg2|time [--variant V] [--mode M]

Every case prints one JSON line; the watchdog word of the scratch is reported
so a protocol error reads as data, not as a hang.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import cadence_gemma_b200 as cg  # noqa: E402
from cadence_gemma_b200 import _abi  # noqa: E402


def make(B, T, E, H, seed=0, resets=True):
  g = torch.Generator().manual_seed(seed)
  bw = E // H
  dev = torch.device("cuda")
  x = torch.randn((B, T, E), generator=g).to(torch.bfloat16).to(dev)
  torch.manual_seed(seed)                  # a_param comes from reset_parameters (global generator)
  lru = cg.RGLRU(E, H, device=dev, dtype=torch.bfloat16)
  with torch.no_grad():
    lru.input_gate.w.copy_((torch.randn((H, bw, bw), generator=g) * bw ** -0.5).to(torch.bfloat16))
    lru.a_gate.w.copy_((torch.randn((H, bw, bw), generator=g) * bw ** -0.5).to(torch.bfloat16))
    lru.input_gate.b.copy_(torch.randn((H, bw), generator=g).to(torch.bfloat16))
    lru.a_gate.b.copy_(torch.randn((H, bw), generator=g).to(torch.bfloat16))
  seg = torch.arange(T, dtype=torch.int32)[None].repeat(B, 1)
  if resets and T > 8:
    for b in range(B):
      cut = int(torch.randint(1, T, (1,), generator=g))
      seg[b, cut:] = torch.arange(T - cut, dtype=torch.int32)
  h0 = torch.randn((B, E), generator=g).to(dev)
  return x, lru, seg.to(dev), h0


def stats(a, b):
  a32, b32 = a.float(), b.float()
  return {"max_abs": (a32 - b32).abs().max().item(),
          "normwise": ((a32 - b32).abs().max() / b32.abs().max().clamp_min(1e-30)).item(),
          "identical": (a == b).float().mean().item()}


def run_case(B, T, E, H, mode, variant, use_h0=True, debug=True):
  x, lru, seg, h0 = make(B, T, E, H)
  wpack = _abi.pack_gate_weights(lru.input_gate.w, lru.a_gate.w)
  ws = _abi.fused_workspace(x.device, B, T, E)
  arith = mode | (variant << 8)
  out = _abi.rglru_fused_fwd(x, wpack, lru.input_gate.b, lru.a_gate.b, lru.a_param, seg, H,
                             h0=h0 if use_h0 else None, arith_mode=arith, debug=debug, workspace=ws)
  torch.cuda.synchronize()
  res = {"case": [B, T, E, H], "mode": mode, "variant": variant}
  y, last_h = out[0], out[1]
  # unfused reference: cuBLAS fused-gate GEMM + scan kernel, same arithmetic mode
  gates = lru.gate_gemm(x)
  y_ref, h_ref = _abi.rglru_fwd(x, None, None, lru.input_gate.b, lru.a_gate.b, lru.a_param, seg,
                                h0=h0 if use_h0 else None, arith_mode=mode, gemm_fused=gates,
                                block_width=E // H)
  torch.cuda.synchronize()
  if debug:
    dbg = out[2]
    g4 = gates.view(B, T, H, 2, E // H)
    res["pre_x"] = stats(dbg[0], g4[:, :, :, 0].reshape(B, T, E))
    res["pre_a"] = stats(dbg[1], g4[:, :, :, 1].reshape(B, T, E))
    res["x_t"] = stats(dbg[2], x)
  res["y"] = stats(y, y_ref)
  csum = lambda t: int(t.contiguous().view(torch.int16).to(torch.int64).sum().item())
  res["checksum"] = {"y_fused": csum(y), "y_unfused": csum(y_ref), "gates_gemm": csum(gates),
                     "h_fused": float(last_h.double().sum().item()), "h_unfused": float(h_ref.double().sum().item())}
  res["last_h"] = stats(last_h, h_ref)
  print(json.dumps(res), flush=True)


def time_case(mode, variant, iters=20, B=8, T=2048):
  E, H = 2560, 10
  x, lru, seg, _ = make(B, T, E, H, resets=False)
  wpack = _abi.pack_gate_weights(lru.input_gate.w, lru.a_gate.w)
  ws = _abi.fused_workspace(x.device, B, T, E)
  y = torch.empty_like(x)
  arith = mode | (variant << 8)

  def call():
    return _abi.rglru_fused_fwd(x, wpack, lru.input_gate.b, lru.a_gate.b, lru.a_param, seg, H,
                                arith_mode=arith, out=y, workspace=ws)

  for _ in range(3):
    call()
  torch.cuda.synchronize()
  evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
  for a, b in evs:
    a.record(); call(); b.record()
  torch.cuda.synchronize()
  us = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
  nelem = B * T * E
  res = {"time": [B, T], "mode": mode, "variant": variant, "us_median": us[len(us) // 2], "us_best": us[0],
         "alg_GBps_8B_per_elem": 8 * nelem / (us[len(us) // 2] * 1e-6) / 1e9}
  print(json.dumps(res), flush=True)


def _time(fn, iters):
  for _ in range(5):
    fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(iters):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) * 1e3 / iters


def conv_case(B=8, T=2048, iters=40, check=True):
  """Conv1D -> RG-LRU: the two kernels one after the other vs ONE fused launch
  (cg_recurrent_prefill_fwd, convolution inside the tcgen05 kernel)."""
  from cadence_gemma_b200 import pipeline
  E, H = 2560, 10
  x, lru, seg, _ = make(B, T, E, H, resets=check)
  conv = cg.Conv1D(E, 4, device=x.device, dtype=torch.bfloat16)
  with torch.no_grad():
    conv.w.normal_(0, 0.4); conv.b.normal_(0, 0.2)
  y = torch.empty_like(x)
  xc = torch.empty_like(x)
  h = torch.empty((B, E), dtype=torch.float32, device=x.device)
  cs = torch.empty((B, 3, E), dtype=x.dtype, device=x.device)

  def run():
    return cg.recurrent_hot_path(conv, lru, x, seg, out=y, last_h_out=h, conv_out=xc, conv_cache_out=cs)

  res = {"conv_case": [B, T]}
  with torch.no_grad():
    old = pipeline.set_fused_conv(False)
    res["two_kernels_us"] = _time(run, iters)
    y1, c1, h1 = [t.clone() for t in run()]
    y.fill_(float("nan")); h.fill_(float("nan")); cs.fill_(float("nan"))   # nothing stale may pass as the second route's output
    pipeline.set_fused_conv(True)
    assert pipeline.can_fuse_conv(conv, lru, x)
    res["one_launch_us"] = _time(run, iters)
    y2, c2, h2 = [t.clone() for t in run()]
    pipeline.set_fused_conv(old)
    res["identical"] = bool(torch.equal(y1, y2) and torch.equal(h1, h2) and torch.equal(c1, c2))
    if not res["identical"]:
      res["y"] = stats(y2, y1); res["h"] = stats(h2, h1); res["cache_equal"] = bool(torch.equal(c1, c2))
  print(json.dumps(res), flush=True)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--case", default="small")
  ap.add_argument("--mode", type=int, default=2)
  ap.add_argument("--variant", type=int, default=0)
  a = ap.parse_args()
  if a.case == "small":
    run_case(1, 64, 256, 1, a.mode, a.variant)          # one family pair, one tile
    run_case(2, 128, 512, 2, a.mode, a.variant)         # look-back across two tiles
  elif a.case == "ragged":
    run_case(3, 200, 512, 2, a.mode, a.variant)         # T not a multiple of the tile
    run_case(2, 33, 256, 2, a.mode, a.variant, use_h0=False)   # head width 128
    run_case(1, 1000, 2560, 10, a.mode, a.variant)
  elif a.case == "cfg2":
    run_case(8, 2048, 2560, 10, a.mode, a.variant, debug=False)
  elif a.case == "time":
    time_case(a.mode, a.variant)
  elif a.case == "conv":
    conv_case()
    conv_case(check=False)
  elif a.case == "conv_shapes":
    for B, T in [(1, 2048), (2, 2048), (4, 2048), (1, 8192), (8, 512), (8, 1024), (16, 256), (2, 8192), (32, 768)]:
      conv_case(B, T, check=False)
  elif a.case == "time_shapes":
    for B, T in [(1, 2048), (2, 2048), (2, 8192), (32, 768), (8, 2048)]:
      time_case(a.mode, a.variant, B=B, T=T)


if __name__ == "__main__":
  main()
