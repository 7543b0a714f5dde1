"""Conv1D / fused RG-LRU kernel times over batch sizes (small-batch prefill regime).

    python scripts/time_small_batch.py        # one JSON line per shape
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cadence_gemma_b200 as cg  # noqa: E402


def timed(fn, iters=50):
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


def main():
  dev = torch.device("cuda")
  E, H = 2560, 10
  torch.manual_seed(0)
  conv = cg.Conv1D(E, 4, device=dev, dtype=torch.bfloat16)
  lru = cg.RGLRU(E, H, device=dev, dtype=torch.bfloat16)
  with torch.no_grad():
    conv.w.normal_(0, 0.4)
    for B, T in [(1, 768), (1, 2048), (2, 2048), (4, 2048), (8, 2048), (1, 8192), (2, 8192)]:
      x = torch.randn((B, T, E), device=dev).to(torch.bfloat16)
      seg = torch.arange(T, dtype=torch.int32, device=dev)[None].repeat(B, 1)
      xc = torch.empty_like(x)
      y = torch.empty_like(x)
      cache = torch.empty((B, 3, E), dtype=torch.bfloat16, device=dev)
      h = torch.empty((B, E), dtype=torch.float32, device=dev)
      t_conv = timed(lambda: conv.forward_into(x, seg, out=xc, cache_out=cache))
      t_lru = timed(lambda: lru.forward_into(xc, seg, out=y, last_h_out=h))
      t_both = timed(lambda: (conv.forward_into(x, seg, out=xc, cache_out=cache),
                              lru.forward_into(xc, seg, out=y, last_h_out=h)))
      t_hot = timed(lambda: cg.recurrent_hot_path(conv, lru, x, seg, out=y, last_h_out=h, conv_cache_out=cache))
      gh = cg.GraphedHotPath(conv, lru, B, T)
      gh.x.copy_(x)
      t_graph = timed(gh.replay)
      print(json.dumps({"B": B, "T": T, "us_conv1d": round(t_conv, 1), "us_rglru_fused": round(t_lru, 1),
                        "us_two_kernels_eager": round(t_both, 1), "us_hot_path_eager": round(t_hot, 1),
                        "us_hot_path_graphed": round(t_graph, 1),
                        "tokens_per_s_M_graphed": round(B * T / t_graph, 1)}), flush=True)


if __name__ == "__main__":
  main()
