"""BASELINE config 5 as microseconds per MODEL decode step: the 18 recurrent blocks of
RecurrentGemma-2B (linear_y, linear_x, Conv1D step, gate GEMVs, RG-LRU step, gating product,
linear_out each), batch 32 (one GPU's slice of the batch of 256) and batch 256, caches warmed
by a T = 16 prefill: eager modules (one-launch fused decode kernel / three small kernels)
against the graphed step (cadence_gemma_b200.decode.GraphedRecurrentDecode).

    python scripts/bench_decode.py > gpurun_out/r2_decode.json
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cadence_gemma_b200 import pipeline  # noqa: E402
from cadence_gemma_b200.decode import GraphedRecurrentDecode  # noqa: E402
from cadence_gemma_b200.modules import RecurrentBlock  # noqa: E402

DEV = torch.device("cuda", 0)


def timeit(fn, iters):
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
  width, heads, nblocks = 2560, 10, 18
  torch.manual_seed(4)
  blocks = [RecurrentBlock(width=width, num_heads=heads, lru_width=width, device=DEV, dtype=torch.bfloat16).eval()
            for _ in range(nblocks)]
  out = {"what": "18 recurrent blocks of RecurrentGemma-2B, one decode step (T = 1), bf16; us per MODEL step",
         "rows": []}
  with torch.no_grad():
    for bsz in (32, 256):
      x0 = torch.randn(bsz, 16, width, device=DEV).to(torch.bfloat16)
      seg = torch.arange(16, device=DEV, dtype=torch.int32)[None].repeat(bsz, 1)
      caches, x = [], x0
      for blk in blocks:
        x, c = blk(x, seg)
        caches.append(c)
      xt = torch.randn(bsz, 1, width, device=DEV).to(torch.bfloat16)
      pos = torch.full((bsz, 1), 16, device=DEV, dtype=torch.int32)

      def eager():
        h = xt
        for blk, c in zip(blocks, caches):
          h, _ = blk(h, pos, c)
        return h

      row = {"batch": bsz}
      old = pipeline.set_fused_decode(True)
      row["eager_one_launch_kernel_us"] = timeit(eager, 30)
      pipeline.set_fused_decode(False)
      row["eager_three_kernels_us"] = timeit(eager, 30)
      pipeline.set_fused_decode(old)
      dec = GraphedRecurrentDecode(blocks, caches, pos)
      row["graphed_us"] = timeit(lambda: dec.step(xt), 100)
      row["graphed_us_per_block"] = row["graphed_us"] / nblocks
      row["tokens_per_s_graphed"] = bsz / (row["graphed_us"] * 1e-6)
      out["rows"].append(row)
  print(json.dumps(out))


if __name__ == "__main__":
  main()
