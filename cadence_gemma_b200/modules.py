"""Caller-side mirror of the reference's ``RecurrentBlock`` (modules.py:503-685).

The reference block is kept as the boundary: two input projections, the
temporal Conv1D, the RG-LRU, the gating product and the output projection.
Only ``conv_1d`` and ``rg_lru`` run custom kernels; the three ``nn.Linear``
stay cuBLAS GEMMs.  On the fused tensor-core path the gating product ``x * y``
(reference modules.py:651) is folded into the RG-LRU kernel's store.  Parameter names / state-dict keys match the reference
(``linear_x``, ``linear_y``, ``linear_out``, ``conv_1d.{w,b}``,
``rg_lru.{a_param,input_gate.{w,b},a_gate.{w,b}}``), so reference checkpoints
load with ``load_state_dict``.
"""
from __future__ import annotations

import math
from typing import NamedTuple

import torch
from torch import nn

from cadence_gemma_b200 import layers, pipeline


class RecurrentBlockCache(NamedTuple):
  """The cache of a recurrent block (reference modules.py:33-38)."""
  rg_lru_state: torch.Tensor    # [B, lru_width] fp32
  conv1d_state: torch.Tensor    # [B, temporal_width - 1, lru_width]


class RecurrentBlock(nn.Module):
  """Griffin / Hawk recurrent block (reference modules.py:503-685)."""

  def __init__(self, width, num_heads, lru_width=None, conv1d_temporal_width=4,
               final_w_init_variance_scale=1.0, device=None, dtype=None):
    super().__init__()
    self.width = width
    self.num_heads = num_heads
    self.lru_width = lru_width or width
    self.conv1d_temporal_width = conv1d_temporal_width
    self.final_w_init_variance_scale = final_w_init_variance_scale
    kw = dict(device=device, dtype=dtype)
    self.linear_y = nn.Linear(self.width, self.lru_width, **kw)
    self.linear_x = nn.Linear(self.width, self.lru_width, **kw)
    self.linear_out = nn.Linear(self.lru_width, self.width, **kw)
    self.conv_1d = layers.Conv1D(self.lru_width, self.conv1d_temporal_width, **kw)
    self.rg_lru = layers.RGLRU(self.lru_width, self.num_heads, **kw)
    self.reset_parameters()

  def reset_parameters(self) -> None:
    in_std = math.sqrt(1.0 / self.width)
    out_std = math.sqrt(self.final_w_init_variance_scale / self.lru_width)
    for lin, std in ((self.linear_x, in_std), (self.linear_y, in_std),
                     (self.linear_out, out_std)):
      nn.init.normal_(lin.weight, mean=0.0, std=std)
      nn.init.zeros_(lin.bias)
    self.conv_1d.reset_parameters()
    self.rg_lru.reset_parameters()

  def forward(self, x, segment_pos, cache=None, return_cache=True):
    """Returns ``(out, RecurrentBlockCache | None)`` (reference :613-660)."""
    gate = self.linear_y(x)              # y branch (no GELU in this fork, :634)
    h = self.linear_x(x)                 # x branch
    # Conv1D -> RG-LRU through the hot-path entry point (pipeline.py).  The gating
    # product (reference :651) rides on the one-launch decode step; folding it
    # into the prefill kernel's store is optional (layers.set_fold_gate)
    fold = (layers.fold_gate_enabled() and self.rg_lru.uses_fused_kernel(h)) or \
        pipeline.can_fuse_decode(self.conv_1d, self.rg_lru, h,
                                 None if cache is None else cache.conv1d_state)
    h, conv_state, lru_state = pipeline.recurrent_hot_path(
        self.conv_1d, self.rg_lru, h, segment_pos,
        conv_cache=None if cache is None else cache.conv1d_state,
        lru_cache=None if cache is None else cache.rg_lru_state,
        return_cache=return_cache, gate_mul=gate if fold else None)
    out = self.linear_out(h if fold else h * gate)
    if not return_cache:
      return out, None
    return out, RecurrentBlockCache(rg_lru_state=lru_state, conv1d_state=conv_state)

  @classmethod
  def init_cache(cls, batch_size, lru_width, dtype, conv1d_temporal_width=4,
                 device=None) -> RecurrentBlockCache:
    return RecurrentBlockCache(
        rg_lru_state=layers.RGLRU.init_cache(batch_size, lru_width, device),
        conv1d_state=layers.Conv1D.init_cache(
            batch_size=batch_size, width=lru_width, dtype=dtype,
            conv1d_temporal_width=conv1d_temporal_width, device=device))
