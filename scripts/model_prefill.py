"""BASELINE.json configs[2] at MODEL level: Cadence multimodal prefill.

    python bench.py --workload config3 [--gpus N]        (driver-compatible JSON line)
    python bench.py --workload config3 --impl reference  (the reference's own classes on the host cores)

What is measured: synthetic ViT features
``[32, 256, 2176]`` (the timm SigLIP + DINOv2 encoder cannot be built offline)
-> ``MLPProjector`` (reference projector/mlp.py:7-30) -> ``[32, 256, 2560]``
visual tokens prepended to 512 embedded text tokens (reference
torch/griffin.py:174-191) -> the RecurrentGemma-2B block stack (common.py:90-101:
width 2560, 10 heads, 26 blocks = (R, R, A) x 8 + (R, R), MLP 7680, local
attention window 2048) with ``segment_pos = cat(arange(256), arange(512))`` --
document starts at 0 and 256 -- returning every block's cache, then the final
norm and the logits of the LAST position (what the sampler's prefill needs,
examples/cadence_sampler.py:229-243).  ``Griffin.forward`` itself cannot take a
batch (quirk D4, griffin.py:171-172), so the forward is restated here; the
reference classes it restates are cited at each piece.

The 18 recurrent blocks run on the sm_100a kernels through
``cadence_gemma_b200.modules.RecurrentBlock`` (hot path =
``pipeline.recurrent_hot_path``); attention, MLP, norms and the dense linears
are plain torch (cuBLAS / SDPA) and OUT OF SCOPE of this repo -- they are here so
that the hot path is measured inside the model it belongs to, with its share of
the time reported.  Batch-sharded over the ranks (strong scaling: the batch of
32 is split, B/N rows per GPU), no collective inside the forward, ONE NCCL
all-gather of the final-token logits.
"""
from __future__ import annotations

import json
import math
import os
import statistics
import sys
import time

import torch
from torch import nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

CFG = dict(vocab=256_000, width=2560, mlp=7680, heads=10, lru_width=2560, blocks=26, window=2048,
           soft_cap=30.0, batch=32, visual=256, text=512, feat=2176)
METRIC = "cadence_multimodal_prefill_tokens_per_sec"
UNIT = "tokens/s"
_MIN_LOGIT = -2.3819763e38


def block_types():
  return ["R", "R", "A"] * 8 + ["R", "R"]          # common.py:56-60, :96


class RMSNorm(nn.Module):                          # reference layers.py:35-79
  def __init__(self, width, device, dtype):
    super().__init__()
    self.scale = nn.Parameter(torch.zeros(width, device=device, dtype=dtype))
    self.eps = 1e-6

  def forward(self, x):
    var = torch.mean(torch.square(x), dim=-1, keepdim=True)
    return x * torch.rsqrt(var + self.eps) * (self.scale + 1)


class MLPBlock(nn.Module):                         # reference modules.py:688-757 (gated GeLU)
  def __init__(self, width, expanded, device, dtype):
    super().__init__()
    self.up = nn.Parameter(torch.randn(2, width, expanded, device=device, dtype=dtype) * width ** -0.5)
    self.up_b = nn.Parameter(torch.zeros(2, 1, 1, expanded, device=device, dtype=dtype))
    self.down = nn.Linear(expanded, width, device=device, dtype=dtype)

  def forward(self, x):
    out = torch.einsum("...td,cdD->c...tD", x, self.up) + self.up_b
    return self.down(F.gelu(out[0], approximate="tanh") * out[1])


def apply_rope(x, positions, max_wavelength=10_000):   # reference modules.py:52-86
  x_rope, x_pass = torch.chunk(x, 2, dim=-1)
  pos = positions[:, :, None, None]
  freq = torch.arange(x_rope.shape[-1] // 2, device=x.device)
  inv = 1.0 / (max_wavelength ** (2 * freq / x_rope.shape[-1]))
  s = pos * inv
  sin, cos = torch.sin(s).type_as(x), torch.cos(s).type_as(x)
  a, b = torch.chunk(x_rope, 2, dim=-1)
  return torch.cat([a * cos - b * sin, b * cos + a * sin, x_pass], dim=-1)


class LocalAttentionBlock(nn.Module):              # reference modules.py:300-500 (MQA, local window)
  def __init__(self, width, heads, window, device, dtype):
    super().__init__()
    self.heads, self.hd, self.window = heads, width // heads, window
    kw = dict(device=device, dtype=dtype)
    self.proj_q = nn.Linear(width, width, bias=False, **kw)
    self.proj_k = nn.Linear(width, self.hd, bias=False, **kw)
    self.proj_v = nn.Linear(width, self.hd, bias=False, **kw)
    self.proj_final = nn.Linear(width, width, **kw)

  def forward(self, x, segment_pos, mask):
    b, t, _ = x.shape
    q = apply_rope(self.proj_q(x).view(b, t, self.heads, self.hd), segment_pos)
    k = apply_rope(self.proj_k(x).view(b, t, 1, self.hd), segment_pos)
    v = self.proj_v(x).view(b, t, 1, self.hd)
    # one K/V head for all query heads (:441-442); mask = causal & window & same document (:88-130)
    o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2).expand(b, self.heads, t, self.hd),
                                       v.transpose(1, 2).expand(b, self.heads, t, self.hd), attn_mask=mask)
    w = min(self.window, t)   # the KV cache the prefill returns (:256-281), unrolled layout
    cache = (k[:, -w:], v[:, -w:], segment_pos[:, -1] + 1)
    return self.proj_final(o.transpose(1, 2).reshape(b, t, -1)), cache


def forward_pass_mask(segment_pos, window):         # reference modules.py:132-155
  seg_ids = torch.cumsum(segment_pos == 0, dim=-1)
  t = segment_pos.shape[-1]
  pos = torch.arange(t, device=segment_pos.device)
  m = (pos[:, None] >= pos[None, :]) & (pos[:, None] <= pos[None, :] + window)
  return (m[None] & (seg_ids[:, :, None] == seg_ids[:, None, :]))[:, None]


class ResidualBlock(nn.Module):                     # reference modules.py:759-914
  def __init__(self, kind, cfg, device, dtype):
    super().__init__()
    from cadence_gemma_b200.modules import RecurrentBlock
    self.kind = kind
    self.temporal_pre_norm = RMSNorm(cfg["width"], device, dtype)
    if kind == "R":
      self.temporal = RecurrentBlock(width=cfg["width"], num_heads=cfg["heads"], lru_width=cfg["lru_width"],
                                     device=device, dtype=dtype)
    else:
      self.temporal = LocalAttentionBlock(cfg["width"], cfg["heads"], cfg["window"], device, dtype)
    self.channel_pre_norm = RMSNorm(cfg["width"], device, dtype)
    self.mlp = MLPBlock(cfg["width"], cfg["mlp"], device, dtype)

  def forward(self, x, segment_pos, mask):
    h = self.temporal_pre_norm(x)
    if self.kind == "R":
      h, cache = self.temporal(h, segment_pos)
    else:
      h, cache = self.temporal(h, segment_pos, mask)
    residual = h + x
    return self.mlp(self.channel_pre_norm(residual)) + residual, cache


class CadencePrefill(nn.Module):
  """projector + embedder + 26 blocks + final norm + last-position logits."""

  def __init__(self, cfg, device, dtype=torch.bfloat16):
    super().__init__()
    self.cfg = cfg
    kw = dict(device=device, dtype=dtype)
    self.projector = nn.Sequential(nn.Linear(cfg["feat"], cfg["width"], **kw), nn.GELU(),      # projector/mlp.py:13-27
                                   nn.Linear(cfg["width"], cfg["width"], **kw), nn.GELU(),
                                   nn.Linear(cfg["width"], cfg["width"], **kw))
    self.embedding = nn.Parameter(torch.randn(cfg["vocab"], cfg["width"], **kw) * cfg["width"] ** -0.5)
    self.blocks = nn.ModuleList(ResidualBlock(k, cfg, device, dtype) for k in block_types()[:cfg["blocks"]])
    self.final_norm = RMSNorm(cfg["width"], device, dtype)
    with torch.no_grad():   # non-trivial gates / conv taps (the reference zero-inits the biases)
      for blk in self.blocks:
        if blk.kind == "R":
          rb = blk.temporal
          rb.rg_lru.input_gate.b.normal_(0, 1); rb.rg_lru.a_gate.b.normal_(0, 1)
          rb.conv_1d.w.normal_(0, 0.3); rb.conv_1d.b.normal_(0, 0.1)

  def forward(self, features, tokens):
    cfg = self.cfg
    b = tokens.shape[0]
    visual = self.projector(features)                                             # griffin.py:181-183
    text = self.embedding[tokens] * torch.tensor(math.sqrt(cfg["width"])).type(torch.bfloat16)   # modules.py:995-1001
    x = torch.cat([visual, text], dim=1)                                          # griffin.py:184
    seg = torch.cat([torch.arange(visual.shape[1]), torch.arange(text.shape[1])]).to(torch.int32)
    seg = seg.to(x.device)[None].repeat(b, 1)                                     # griffin.py:186-191 (resets at 0 and 256)
    mask = forward_pass_mask(seg, cfg["window"])
    caches = []
    for blk in self.blocks:                                                       # griffin.py:196-211
      x, cache = blk(x, seg, mask)
      caches.append(cache)
    last = self.final_norm(x[:, -1:])                                             # griffin.py:215-221, last position
    logits = last @ self.embedding.T
    c = cfg["soft_cap"]
    return torch.tanh(logits / c) * c, caches


# ----------------------------------------------------------------------------------------------
def _config(world):
  c = CFG
  return {"workload": ("BASELINE configs[2]: Cadence ViT+projector+RecurrentGemma-2B multimodal prefill, synthetic ViT "
                       f"features [{c['batch']},{c['visual']},{c['feat']}] (timm encoder unavailable offline) + "
                       f"{c['text']} text tokens, batch {c['batch']}, bf16, random init"),
          "global_batch": c["batch"], "batch_per_gpu": c["batch"] // world, "seq_len": c["visual"] + c["text"],
          "blocks": "18 recurrent (sm_100a kernels) + 8 local-attention (torch SDPA), MLP 7680, vocab 256000",
          "parallelism": f"batch-sharded x{world} (strong scaling), no collective in the forward; one NCCL all-gather of the final-token logits",
          "l2": "activations 126 MB per tensor per 32 rows; weights 5.4 GB: inputs larger than L2"}


def run(args, dev, world, rank):
  import torch.distributed as dist
  import cadence_gemma_b200 as cg
  from cadence_gemma_b200 import _abi, pipeline
  c = CFG
  assert c["batch"] % world == 0
  rows = c["batch"] // world
  torch.manual_seed(2)                       # seed 2; identical weights on every rank
  model = CadencePrefill(c, dev).eval()
  g = torch.Generator().manual_seed(200 + rank)
  feats_h = torch.randn((rows, c["visual"], c["feat"]), generator=g).to(torch.bfloat16).pin_memory()
  toks_h = torch.randint(0, c["vocab"], (rows, c["text"]), generator=g).pin_memory()
  feats, toks = feats_h.to(dev), toks_h.to(dev)
  tokens_total = c["batch"] * (c["visual"] + c["text"])

  # time of the hot path inside the model: events around every recurrent_hot_path call
  hot_events = []
  orig = pipeline.recurrent_hot_path
  record = [False]

  def timed_hot_path(*a, **k):
    if not record[0]:
      return orig(*a, **k)
    e0 = torch.cuda.Event(enable_timing=True); e0.record()
    out = orig(*a, **k)
    e1 = torch.cuda.Event(enable_timing=True); e1.record()
    hot_events.append((e0, e1))
    return out

  pipeline.recurrent_hot_path = timed_hot_path
  import cadence_gemma_b200.modules as cgm
  cgm.pipeline.recurrent_hot_path = timed_hot_path
  gathered = torch.empty((c["batch"], 1, c["vocab"]), dtype=torch.bfloat16, device=dev) if world > 1 else None

  def step(fe, to):
    logits, caches = model(fe, to)
    if world > 1:
      dist.all_gather_into_tensor(gathered, logits.contiguous())
    return logits

  def sync():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
      torch.cuda.synchronize()

  steps = max(1, min(args.steps, 10))
  with torch.no_grad():
    for _ in range(max(args.warmup, 3)):
      out = step(feats, toks)
    sync()
    launches0 = _abi.launch_count
    record[0] = True
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
      out = step(feats, toks)
    t1.record()
    sync()
    record[0] = False
    launches = _abi.launch_count - launches0
    ms = t0.elapsed_time(t1)
    hot_ms = sum(a.elapsed_time(b) for a, b in hot_events)
    # e2e: inputs from pinned host memory, final-token logits back to the host, every step
    logits_h = torch.empty((rows, 1, c["vocab"]), dtype=torch.bfloat16).pin_memory()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for _ in range(steps):
      fe = feats_h.to(dev, non_blocking=True)
      to = toks_h.to(dev, non_blocking=True)
      logits_h.copy_(step(fe, to), non_blocking=True)
    e1.record()
    sync()
    e2e_ms = e0.elapsed_time(e1)
  pipeline.recurrent_hot_path = orig
  cgm.pipeline.recurrent_hot_path = orig

  t = torch.tensor([ms, hot_ms, e2e_ms], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  ms, hot_ms, e2e_ms = t.tolist()
  if rank != 0:
    return None
  from bench import measured_peaks
  peak, peak_tf, peak_src = measured_peaks()
  nelem = rows * (c["visual"] + c["text"]) * c["lru_width"] * 18      # hot-path elements per rank per step
  hot_us = 1e3 * hot_ms / steps
  return {
      "metric": METRIC, "value": tokens_total * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
      "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / steps, "higher_is_better": True,
      "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
      "config": _config(world), "device": "cuda",
      "e2e": {"value": tokens_total * steps / (e2e_ms * 1e-3), "unit": UNIT,
              "h2d_bytes_per_step": feats_h.numel() * 2 + toks_h.numel() * 8,
              "d2h_bytes_per_step": rows * c["vocab"] * 2, "ms_per_step": e2e_ms / steps},
      "gpu_launches": launches,
      "recurrent_hot_path": {
          "what": "Conv1D -> RG-LRU of the 18 recurrent blocks (CUDA events around every recurrent_hot_path call)",
          "ms_per_step": hot_ms / steps, "share_of_step": hot_ms / ms,
          "us_per_block": hot_us / 18},
      "roofline": {"bound": "hbm", "kernel": "Conv1D + fused tcgen05 RG-LRU, 18 blocks, canonical 12 B/element (SURVEY 8d)",
                   "achieved": 12 * nelem / (hot_us * 1e-6) / 1e9, "peak": peak, "unit": "GB/s",
                   "frac": 12 * nelem / (hot_us * 1e-6) / 1e9 / peak, "peak_source": peak_src, "traffic": None},
      "arith_mode": cg.get_arith_mode(),
  }


def reference_arm_line(args):
  """The reference's OWN classes on the host cores: MLPProjector, Embedder, ResidualBlock x 26, final
  norm (baseline/_ref), driven by the same restated forward; a bounded sample of the workload (one
  batch row of the 32: the rows are independent), tokens/s scaled from that sample."""
  from oracle import ref_loader
  c = CFG
  base = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
          "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
          "dtype": "bf16", "data": "synthetic", "config": _config(args.gpus), "device": "cpu", "gpu_launches": 0}
  if not ref_loader.reference_available():
    return {"impl": "reference", "unavailable": "no copy of the reference on this box (baseline/_ref missing)"}
  ref = ref_loader.load_reference()
  threads = os.cpu_count() or 1
  torch.set_num_threads(threads)
  torch.manual_seed(2)
  R, A = ref.common.TemporalBlockType.RECURRENT, ref.common.TemporalBlockType.ATTENTION
  dtype = torch.bfloat16
  blocks = [ref.modules.ResidualBlock(width=c["width"], mlp_expanded_width=c["mlp"], num_heads=c["heads"],
                                      attention_window_size=c["window"],
                                      temporal_block_type=R if k == "R" else A, lru_width=c["lru_width"],
                                      dtype=dtype) for k in block_types()]
  embedder = ref.modules.Embedder(vocab_size=c["vocab"], embed_dim=c["width"], scale_by_sqrt_dim=True, dtype=dtype)
  final_norm = ref.layers.RMSNorm(width=c["width"], dtype=dtype)
  import importlib
  projector = importlib.import_module("recurrentgemma.projector.mlp").MLPProjector(device="cpu")
  g = torch.Generator().manual_seed(200)
  rows = 1
  feats = torch.randn((rows, c["visual"], c["feat"]), generator=g).to(dtype)
  toks = torch.randint(0, c["vocab"], (rows, c["text"]), generator=g)
  seg = torch.cat([torch.arange(c["visual"]), torch.arange(c["text"])])[None].repeat(rows, 1)

  def step():
    with torch.no_grad():
      x = torch.cat([projector(feats), embedder.encode(toks)], dim=1)
      for blk in blocks:
        x, _ = blk(x, seg, None, True)
      logits = embedder.decode(final_norm(x[:, -1:]))
      return torch.tanh(logits / c["soft_cap"]) * c["soft_cap"]

  times = []
  for i in range(args.warmup + args.steps):
    t0 = time.perf_counter()
    step()
    if i >= args.warmup:
      times.append(time.perf_counter() - t0)
  per_step = statistics.mean(times)
  value = rows * (c["visual"] + c["text"]) / per_step
  base.update({"value": value, "ms_per_step": per_step * 1e3 * (c["batch"] / rows),
               "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
                                "sample": f"{rows} of the {c['batch']} batch rows per step (rows are independent), "
                                          f"{len(times)} steps after {args.warmup} warm-ups; ms_per_step scaled to the full batch"},
               "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
  return base


if __name__ == "__main__":
  print(json.dumps(_config(1), indent=1))
