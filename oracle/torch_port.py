"""ORACLE (test infrastructure) -- CPU/eager restatement of the reference path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this file.  The product path
(``cadence_gemma_b200``) never does.

This is a *functional* restatement, in plain eager ``torch`` ops, of the
reference's recurrent hot path (all paths relative to ``/root/reference``):

  * ``rnn_scan``               recurrentgemma/torch/layers.py:146-199
  * ``block_diagonal_linear``  recurrentgemma/torch/layers.py:133-142
  * ``rglru_forward``          recurrentgemma/torch/layers.py:322-375
  * ``conv1d_forward``         recurrentgemma/torch/layers.py:458-546,
                               mask quirk :592-633, padding :635-662
  * ``recurrent_block_forward`` recurrentgemma/torch/modules.py:613-660

It issues the same ATen primitives in the same order as the reference, so on
the same torch build it reproduces the reference bit for bit, including every
bf16 rounding point (each eager op rounds to the tensor dtype).  Parity is
PINNED: ``tests/test_oracle_golden.py`` checks it against golden vectors
produced by the unmodified reference (``tests/golden/make_golden.py``) and,
when ``/root/reference`` is present, against the reference executed live.

It is also the "port" timed as the CPU baseline: it keeps the reference's
cost structure (a Python loop over time steps, one full-tensor eager op per
arithmetic step).
"""
from __future__ import annotations

from typing import NamedTuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# Linear recurrence -- layers.py:146-199
# --------------------------------------------------------------------------
def rnn_scan(x, a, reset, h0, acc_dtype=torch.float32):
  """h_t = a_t * h_{t-1} + x_t along dim 1, accumulated in ``acc_dtype``."""
  assert x.ndim == 3                                   # :166
  assert a.shape == x.shape[-a.ndim:]                  # :167
  assert a.dtype == x.dtype                            # :168
  assert h0 is None or h0.dtype == acc_dtype           # :170
  io_dtype = x.dtype
  keep = ~reset[..., None]
  a = a * keep                                         # :173 (reset zeroes a)
  steps = x.shape[1]
  if steps == 1:                                       # :175-182
    if h0 is None:
      return x, x[:, 0].to(acc_dtype)
    h = a.to(acc_dtype) * h0[:, None] + x.to(acc_dtype)
    return h.to(io_dtype), h[:, -1]
  h = h0 if h0 is not None else torch.zeros(
      x[:, 0].shape, dtype=acc_dtype, device=x.device)  # :187-190
  out = torch.zeros_like(x)                            # :192
  a_acc = a.to(acc_dtype)
  x_acc = x.to(acc_dtype)
  for t in range(steps):                               # :195-197
    h = a_acc[:, t] * h + x_acc[:, t]                  # mul, then add (no FMA)
    out[:, t] = h                                      # cast to io dtype on store
  return out, h


# --------------------------------------------------------------------------
# Block-diagonal gate projection -- layers.py:133-142
# --------------------------------------------------------------------------
def block_diagonal_linear(x, w, b):
  """x[..., H*bw] @ blockdiag(w[H,bw,bw]) + b[H,bw]."""
  heads, bw, _ = w.shape
  xs = x.reshape(*x.shape[:-1], heads, bw)
  y = torch.einsum("...hi,hij->...hj", xs, w) + b       # :139
  return y.reshape(*x.shape[:-1], heads * bw)


class RGLRUParams(NamedTuple):
  a_param: torch.Tensor         # [E]
  input_gate_w: torch.Tensor    # [H,bw,bw]
  input_gate_b: torch.Tensor    # [H,bw]
  a_gate_w: torch.Tensor        # [H,bw,bw]
  a_gate_b: torch.Tensor        # [H,bw]


def rglru_gate_math(x, pre_x, pre_a, a_param, reset):
  """layers.py:348-365 from gate pre-activations to (normalized_x, a)."""
  gate_x = torch.sigmoid(pre_x)                        # :348
  gate_a = torch.sigmoid(pre_a)                        # :349
  log_a = -8.0 * gate_a * F.softplus(a_param)          # :352
  a = torch.exp(log_a)                                 # :353
  a_square = torch.exp(2 * log_a)                      # :354
  gated_x = x * gate_x                                 # :357
  multiplier = torch.sqrt(1 - a_square)                # :361 (fwd of SqrtBound)
  multiplier = reset[..., None] + ~reset[..., None] * multiplier   # :364
  normalized_x = gated_x * multiplier.to(x.dtype)      # :365
  return normalized_x, a


def rglru_forward(p: RGLRUParams, x, segment_pos, cache=None,
                  return_cache=True):
  """layers.py:322-375."""
  bs, length, _ = x.shape
  if segment_pos.shape != (bs, length):                # :342-343
    segment_pos = segment_pos[None, :]
  assert segment_pos.shape == (bs, length)             # :344
  reset = segment_pos == 0                             # :345
  pre_x = block_diagonal_linear(x, p.input_gate_w, p.input_gate_b)
  pre_a = block_diagonal_linear(x, p.a_gate_w, p.a_gate_b)
  normalized_x, a = rglru_gate_math(x, pre_x, pre_a, p.a_param, reset)
  y, last_h = rnn_scan(normalized_x, a, reset, cache)  # :366-371
  return (y, last_h) if return_cache else (y, None)


def rglru_from_preacts(x, pre_x, pre_a, a_param, segment_pos, cache=None):
  """Kernel-boundary oracle: same as ``rglru_forward`` after the two GEMMs.

  ``pre_x`` / ``pre_a`` are the block-diagonal outputs *including* bias, in
  x's dtype -- exactly what ``cg_rglru_fwd`` consumes.
  """
  bs, length, _ = x.shape
  if segment_pos.shape != (bs, length):
    segment_pos = segment_pos[None, :]
  reset = segment_pos == 0
  normalized_x, a = rglru_gate_math(x, pre_x, pre_a, a_param, reset)
  return rnn_scan(normalized_x, a, reset, cache)


# --------------------------------------------------------------------------
# Temporal Conv1D -- layers.py:458-546
# --------------------------------------------------------------------------
def _document_mask(segment_pos, start, end, look_ahead):
  """layers.py:592-633 *as the fork has it* (quirk D1 in SURVEY.md).

  The loop bound is ``range(1, look_ahead - 1)`` (upstream: ``look_ahead + 1``).
  """
  not_boundary = (segment_pos != 0).to(torch.int32)
  mask = torch.ones((segment_pos.shape[0], end - start),
                    device=segment_pos.device)
  for shift in range(1, look_ahead - 1):               # :629
    mask = mask * not_boundary[:, start + shift:end + shift]
  return mask


def conv1d_forward(w, b, x, segment_pos, cache=None, return_cache=True):
  """layers.py:458-546.  ``w``: [W,E], ``b``: [E].

  Unlike the reference this does NOT mutate the caller's ``x`` (quirk D2,
  ``x_window *= mask`` on a view, :524): it works on a private copy, which
  yields the same outputs (the in-place masks only accumulate, see
  SURVEY.md section 8 row A3) and the same returned cache.
  """
  temporal_width = w.shape[0]
  out_len = x.shape[1]
  if cache is not None:                                # :478-483 decode
    bsz, n_tok, width = x.shape
    assert cache.shape == (bsz, temporal_width - 1, width)   # :565
    assert n_tok == 1                                  # :566
    x = torch.cat([cache.to(x.dtype), x], dim=1)       # :567
    prompt_len = temporal_width - 1
    cache_dtype = cache.dtype
  else:
    x = x.clone()                                      # private copy (see doc)
    prompt_len = 0
    cache_dtype = x.dtype
  acc = 0.0                                            # :492
  taps = min(temporal_width, prompt_len + out_len)     # :496
  for shift in range(taps):                            # :500
    start = max(prompt_len - shift, 0)                 # :588
    end = prompt_len + out_len - shift                 # :589
    window = x[:, start:end]
    if cache is None:                                  # :508
      sp = segment_pos if segment_pos.ndim == 2 else segment_pos[None, :]
      mask = _document_mask(sp, start, end, shift)
      window *= mask[:, :, None].to(x.dtype)           # :524 in place on the copy
    pad = out_len - window.shape[1]                    # :635-648
    if pad:
      window = torch.cat([
          torch.zeros((window.shape[0], pad, window.shape[2]),
                      dtype=window.dtype, device=window.device), window], dim=1)
    acc = acc + window * w[temporal_width - shift - 1][None, None, :]   # :530-533
  acc = acc + b[None, None]                            # :536
  if not return_cache:
    return acc, None
  state = x[:, 1 - temporal_width:].to(cache_dtype)    # :542
  pad = temporal_width - 1 - state.shape[1]            # :650-662
  if pad:
    state = torch.cat([
        torch.zeros((state.shape[0], pad, state.shape[2]), dtype=state.dtype,
                    device=state.device), state], dim=1)
  return acc, state


# --------------------------------------------------------------------------
# Caller -- modules.py:613-660 (RecurrentBlock.forward), no GELU (quirk D3)
# --------------------------------------------------------------------------
class RecurrentBlockParams(NamedTuple):
  linear_y_w: torch.Tensor
  linear_y_b: torch.Tensor
  linear_x_w: torch.Tensor
  linear_x_b: torch.Tensor
  linear_out_w: torch.Tensor
  linear_out_b: torch.Tensor
  conv_w: torch.Tensor
  conv_b: torch.Tensor
  rglru: RGLRUParams


def recurrent_block_forward(p: RecurrentBlockParams, x, segment_pos,
                            cache=None, return_cache=True):
  """Returns (out, (rg_lru_state, conv1d_state) | None)."""
  y = F.linear(x, p.linear_y_w, p.linear_y_b)          # :634
  xb = F.linear(x, p.linear_x_w, p.linear_x_b)         # :637
  xb, conv_state = conv1d_forward(
      p.conv_w, p.conv_b, xb, segment_pos,
      cache=None if cache is None else cache[1], return_cache=return_cache)
  xb, lru_state = rglru_forward(
      p.rglru, xb, segment_pos,
      cache=None if cache is None else cache[0], return_cache=return_cache)
  out = F.linear(xb * y, p.linear_out_w, p.linear_out_b)   # :651-652
  if not return_cache:
    return out, None
  return out, (lru_state, conv_state)


# --------------------------------------------------------------------------
# Parameter initialisation -- layers.py:202-221, :127-130, :432-435
# (used by the synthetic-workload generators of bench.py and the tests so the
# value ranges match a real random-init RecurrentGemma block)
# --------------------------------------------------------------------------
def init_a_param(width, generator, dtype=torch.float32, min_rad=0.9,
                 max_rad=0.999, eps=1e-8):
  u = torch.empty(width, dtype=torch.float32).uniform_(
      min_rad**2 + eps, max_rad**2 + eps, generator=generator)
  a_real_log = 0.5 * torch.log(u)
  return torch.log(torch.exp(-a_real_log) - 1.0).to(dtype)


def init_rglru_params(width, heads, generator, dtype=torch.float32,
                      bias_std=1.0):
  bw = width // heads
  std = (1.0 / bw) ** 0.5

  def normal(shape, s):
    return (torch.randn(shape, generator=generator) * s).to(dtype)

  return RGLRUParams(
      a_param=init_a_param(width, generator, dtype),
      input_gate_w=normal((heads, bw, bw), std),
      input_gate_b=normal((heads, bw), bias_std),
      a_gate_w=normal((heads, bw, bw), std),
      a_gate_b=normal((heads, bw), bias_std),
  )


def init_conv_params(width, temporal_width, generator, dtype=torch.float32,
                     w_scale=0.01, bias_std=0.1):
  std = (w_scale / temporal_width) ** 0.5
  w = (torch.randn((temporal_width, width), generator=generator) * std).to(dtype)
  b = (torch.randn((width,), generator=generator) * bias_std).to(dtype)
  return w, b
