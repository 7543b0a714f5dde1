"""Block-level timings of BASELINE.json configs 2-5 on one GPU (device resident).

    python scripts/bench_configs.py [--out gpurun_out/configs.json]

For every config: Conv1D -> fused gate GEMM -> RG-LRU through the module API,
CUDA events around N back-to-back steps.  Config 5 (decode) is measured eager
and as a CUDA graph of one step.  GB/s are algorithmic bytes (6*s per element
for the two custom kernels) over the custom-kernel time.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cadence_gemma_b200 as cg  # noqa: E402
from cadence_gemma_b200 import _abi  # noqa: E402

DEV = "cuda:0"
E, H, W = 2560, 10, 4


def modules(dtype):
  torch.manual_seed(0)
  conv = cg.Conv1D(E, W, device=DEV, dtype=dtype)
  lru = cg.RGLRU(E, H, device=DEV, dtype=dtype)
  with torch.no_grad():
    conv.w.normal_(0, 0.3); conv.b.normal_(0, 0.1)
    lru.input_gate.b.normal_(); lru.a_gate.b.normal_()
  return conv, lru


def timed(fn, iters):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  s.record()
  for _ in range(iters):
    fn()
  e.record(); e.synchronize()
  return s.elapsed_time(e) * 1e3 / iters


def prefill_case(name, B, T, seg, dtype, iters=20):
  conv, lru = modules(dtype)
  x = torch.randn((B, T, E), device=DEV).to(dtype)
  seg = seg.to(DEV)
  s = x.element_size()
  with torch.no_grad():
    def full():
      xc, _ = conv(x, seg)
      return lru(xc, seg)
    xc, _ = conv(x, seg)
    fused = lru.uses_fused_kernel(xc)
    t_full = timed(full, iters)
    t_conv = timed(lambda: conv(x, seg), iters)
    if fused:   # gate GEMMs + gates + scan in one tcgen05 kernel
      t_gemm = 0.0
      t_lru = timed(lambda: lru(xc, seg), iters)
    else:
      gates = lru.gate_gemm(xc)
      t_gemm = timed(lambda: lru.gate_gemm(xc), iters)
      t_lru = timed(lambda: _abi.rglru_fwd(xc, None, None, lru.input_gate.b, lru.a_gate.b,
                                           lru.a_param, seg, arith_mode=cg.get_arith_mode(),
                                           gemm_fused=gates, block_width=E // H), iters)
    # the same step on the unfused pair of kernels, for comparison
    old = cg.set_fused(False)
    try:
      t_full_unfused = timed(full, iters)
    finally:
      cg.set_fused(old)
  n = B * T * E
  return dict(config=name, B=B, T=T, dtype=str(dtype), fused_tcgen05=bool(fused), us_step=t_full,
              us_step_unfused=t_full_unfused, us_conv1d=t_conv,
              us_gate_gemm=t_gemm, us_rglru=t_lru, tokens_per_s=B * T / t_full * 1e6,
              gbps_conv1d=2 * s * n / t_conv / 1e3, gbps_rglru=4 * s * n / t_lru / 1e3,
              gbps_block=6 * s * n / (t_conv + t_gemm + t_lru) / 1e3)


def block_case(B, T, dtype, iters=20):
  """Whole RecurrentBlock (reference modules.py:613-660) at 2B shapes: three cuBLAS
  projections + the hot path, with / without the fused kernel and its folded
  gating product."""
  from cadence_gemma_b200.modules import RecurrentBlock
  torch.manual_seed(0)
  blk = RecurrentBlock(width=E, num_heads=H, lru_width=E, device=DEV, dtype=dtype)
  with torch.no_grad():
    blk.rg_lru.input_gate.b.normal_(); blk.rg_lru.a_gate.b.normal_()
    x = torch.randn((B, T, E), device=DEV).to(dtype)
    seg = torch.arange(T, dtype=torch.int32, device=DEV)[None].repeat(B, 1)
    run = lambda: blk(x, seg)
    t_nofold = timed(run, iters)      # default: gating product as a separate elementwise kernel
    oldf = cg.set_fold_gate(True)     # optional: folded into the fused kernel's store
    try:
      t_fused = timed(run, iters)
    finally:
      cg.set_fold_gate(oldf)
    old = cg.set_fused(False)
    try:
      t_unfused = timed(run, iters)
    finally:
      cg.set_fused(old)
  return dict(config=f"RecurrentBlock 2B shape B={B} T={T}", dtype=str(dtype),
              us_block_fused_folded=t_fused, us_block_fused_separate_product=t_nofold,
              us_block_unfused=t_unfused, tokens_per_s=B * T / t_fused * 1e6)


def decode_case(B, dtype, iters=200):
  conv, lru = modules(dtype)
  with torch.no_grad():
    xp = torch.randn((B, 16, E), device=DEV).to(dtype)        # warm the caches (T = 16 prefill)
    segp = torch.arange(16, device=DEV)[None].repeat(B, 1)
    xc, conv_cache = conv(xp, segp)
    _, h = lru(xc, segp)
    x = torch.randn((B, 1, E), device=DEV).to(dtype)
    seg = torch.full((B, 1), 16, device=DEV, dtype=torch.int32)
    state = {"c": conv_cache, "h": h}

    def step():   # fused decode step (one launch) where the shape allows it
      y, state["c"], state["h"] = cg.recurrent_hot_path(conv, lru, x, seg, conv_cache=state["c"],
                                                        lru_cache=state["h"])
      return y
    t_eager = timed(step, iters)
    def step3():  # the three-launch path: conv step, cuBLAS GEMV, gate step
      xc1, state["c"] = conv(x, seg, state["c"])
      y, state["h"] = lru(xc1, seg, state["h"])
      return y
    t_eager3 = timed(step3, iters)
    # CUDA graph of one step with static buffers (the ABI is capture-safe)
    sc, sh = conv_cache.clone(), h.clone()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
      for _ in range(3):
        y, c2, h2 = cg.recurrent_hot_path(conv, lru, x, seg, conv_cache=sc, lru_cache=sh)
      torch.cuda.synchronize()
      with torch.cuda.graph(g, stream=side):
        y, c2, h2 = cg.recurrent_hot_path(conv, lru, x, seg, conv_cache=sc, lru_cache=sh)
        sc.copy_(c2); sh.copy_(h2)
      g3 = torch.cuda.CUDAGraph()
      with torch.cuda.graph(g3, stream=side):
        xc1, c2 = conv(x, seg, sc)
        y, h2 = lru(xc1, seg, sh)
        sc.copy_(c2); sh.copy_(h2)
    torch.cuda.current_stream().wait_stream(side)
    t_graph = timed(g.replay, iters)
    t_graph3 = timed(g3.replay, iters)
  best_graph = min(t_graph, t_graph3)
  return dict(config="5 decode step", B=B, T=1, dtype=str(dtype), us_step_eager_one_launch=t_eager,
              us_step_eager_three_launches=t_eager3, us_step_cuda_graph_one_launch=t_graph,
              us_step_cuda_graph_three_launches=t_graph3, tokens_per_s_graph=B / best_graph * 1e6,
              tokens_per_s_eager=B / t_eager * 1e6)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--out", default="gpurun_out/configs.json")
  args = ap.parse_args()
  out = []
  bf = torch.bfloat16
  ar = lambda B, T: torch.arange(T, dtype=torch.int32)[None].repeat(B, 1)
  out.append(prefill_case("2 block bf16 B=8 T=2048", 8, 2048, ar(8, 2048), bf))
  out.append(prefill_case("2 block fp32 B=8 T=2048", 8, 2048, ar(8, 2048), torch.float32))
  seg3 = torch.cat([torch.arange(256), torch.arange(512)]).to(torch.int32)[None].repeat(32, 1)
  out.append(prefill_case("3 multimodal prefill block B=32 T=256+512", 32, 768, seg3, bf))
  g = torch.Generator().manual_seed(3)
  for B in (16, 2):
    seg4 = ar(B, 8192)
    for b in range(B):
      for cut in sorted(torch.randint(1, 8192, (7,), generator=g).tolist()):
        seg4[b, cut:] = torch.arange(8192 - cut, dtype=torch.int32)
    out.append(prefill_case(f"4 long context T=8192 resets B={B}", B, 8192, seg4, bf, iters=8))
  out.append(prefill_case("sampler prefill B=1 T=2048", 1, 2048, ar(1, 2048), bf))
  out.append(block_case(8, 2048, bf))
  out.append(block_case(32, 768, bf))
  out.append(decode_case(32, bf))
  out.append(decode_case(256, bf))
  for r in out:
    print(json.dumps(r), flush=True)
  os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
  json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
  main()
