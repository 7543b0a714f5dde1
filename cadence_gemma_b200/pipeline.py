"""Conv1D -> RG-LRU of one recurrent block: the hot-path entry point.

``recurrent_hot_path(conv, lru, x, segment_pos)`` computes what
``RecurrentBlock.forward`` computes between its projections (reference
``recurrentgemma/torch/modules.py:638-649``):

    x, conv1d_state = conv_1d(x, segment_pos)
    x, rg_lru_state = rg_lru(x, segment_pos)

Prefill at RecurrentGemma shapes (bf16, temporal width 4, head width 128 / 256,
T > 1) can run as ONE kernel: ``cg_recurrent_prefill_fwd`` -- the temporal
convolution runs inside the fused tcgen05 RG-LRU kernel (TMA-loaded rows
convolved in place in shared memory by the epilogue warpgroups), so neither the
conv output nor the gate pre-activations reach HBM.  ``set_fused_conv`` /
``CG_B200_FUSED_CONV`` select it: "auto" (default) takes the one-launch route
whenever the shape allows (it is the faster route at every measured size);
``False`` / "0" runs the Conv1D kernel followed by the fused RG-LRU kernel; other
shapes / fp32
take the Conv1D kernel + cuBLAS gate GEMM + scan kernel.  A decode step (T == 1, caches given) is one launch as well
(``cg_recurrent_decode_step``).

(A producer / consumer overlap of the two kernels on two streams was built and
measured in round 1 -- 287 us vs 137 us at config 2 -- and removed; DESIGN.md
section 9.)

No CPU path.
"""
from __future__ import annotations

import os

import torch

from cadence_gemma_b200 import _abi, layers

# Convolution inside the fused RG-LRU kernel (one launch per prefill step):
# "auto" (default) and "1" = whenever the shape allows, "0" = never (Conv1D kernel, then the fused
# RG-LRU kernel).  The convolution runs in the epilogue warpgroups of the tcgen05 kernel, once per
# head (the two CTAs of a head form a cluster and exchange their halves), and is the faster route at
# every measured shape (profiles/r3_fused_conv_shapes.txt: B=8, T=2048 125 us vs 130 us; B=2, T=8192
# 146 vs 154; B=16, T=8192 873 vs 890; B=32, T=768 177 vs 180; B=1, T=2048 53 vs 79; B=16, T=256
# 42 vs 83).  (Round 2's first version -- two dedicated warps, the whole head per CTA -- lost above
# 128 scan tiles per family: 165-175 us at B=8, T=2048; `FUSED_CONV_MAX_TILES` was its crossover.)
_fused_conv = os.environ.get("CG_B200_FUSED_CONV", "auto")
FUSED_CONV_MAX_TILES = None     # auto: no limit any more (kept for callers that read it)
# One-launch decode step (cg_recurrent_decode_step).  It more than halves the cost
# of an EAGER decode step (30 us vs 76 us per block at B = 32: three launches and
# their host overhead become one); inside a CUDA graph, where launch overhead is
# gone, the three small kernels are ~5 us faster at B = 32 -- switch it off there.
_fused_decode = os.environ.get("CG_B200_FUSED_DECODE", "1") != "0"


def set_fused_conv(mode):
  """Selects the in-kernel convolution of the prefill path: ``"auto"``, ``True`` (always,
  when the shape allows) or ``False`` (never); returns the old setting."""
  global _fused_conv
  old = _fused_conv
  _fused_conv = mode if mode == "auto" else ("1" if mode in (True, "1", 1) else "0")
  return old


def set_fused_decode(enabled: bool) -> bool:
  """Switches the one-launch decode step; returns the old setting."""
  global _fused_decode
  old, _fused_decode = _fused_decode, bool(enabled)
  return old


def can_fuse_decode(conv, lru, x, conv_cache) -> bool:
  """True if a decode step (T == 1, conv cache given) runs as ONE fused launch
  (``cg_recurrent_decode_step``)."""
  return (_fused_decode and layers.fused_enabled() and conv_cache is not None and x.is_cuda and x.shape[1] == 1 and
          (layers.get_arith_mode() & (_abi.ARITH_FP32 | _abi.ARITH_STRICT)) == 0 and
          conv_cache.dtype in (torch.bfloat16, torch.float32) and
          _abi.decode_supported(lru.width, lru.num_heads, conv.w.shape[0], x.dtype))


def can_fuse_conv(conv, lru, x, conv_cache=None) -> bool:
  """True if ``recurrent_hot_path`` runs Conv1D + RG-LRU as ONE kernel
  (``cg_recurrent_prefill_fwd``)."""
  if _fused_conv == "0":
    return False
  if (_fused_conv == "auto" and FUSED_CONV_MAX_TILES is not None and
      x.shape[0] * ((x.shape[1] + 31) // 32) > FUSED_CONV_MAX_TILES):
    return False
  return (layers.fused_enabled() and conv_cache is None and x.is_cuda and
          conv.w.shape[0] == 4 and conv.w.dtype == x.dtype and
          (layers.get_arith_mode() & _abi.ARITH_FP32) == 0 and
          layers.uses_fused_kernel(lru, x))


def recurrent_hot_path(conv, lru, x, segment_pos, conv_cache=None, lru_cache=None,
                       return_cache=True, gate_mul=None, out=None, last_h_out=None,
                       conv_out=None, conv_cache_out=None):
  """``Conv1D.forward -> RGLRU.forward`` of one recurrent block.  With grad enabled
  and anything that requires grad this is the differentiable sequence of the two
  modules (training path); otherwise the inference kernels below."""
  if layers._wants_grad(x, conv_cache, lru_cache, gate_mul, *conv.parameters(), *lru.parameters()):
    assert out is None and last_h_out is None and conv_out is None and conv_cache_out is None
    xc, conv_state = conv(x, segment_pos, conv_cache, return_cache)
    y, h = lru(xc, segment_pos, lru_cache, return_cache)
    return (y if gate_mul is None else y * gate_mul), conv_state, h
  return _recurrent_hot_path_inference(conv, lru, x, segment_pos, conv_cache, lru_cache, return_cache,
                                       gate_mul, out, last_h_out, conv_out, conv_cache_out)


@torch.no_grad()
def _recurrent_hot_path_inference(conv, lru, x, segment_pos, conv_cache=None, lru_cache=None,
                                  return_cache=True, gate_mul=None, out=None, last_h_out=None,
                                  conv_out=None, conv_cache_out=None):
  """Returns ``(y, conv1d_state | None, rg_lru_state | None)``.

  ``conv`` / ``lru``: ``Conv1D`` / ``RGLRU`` modules (ours or the reference's
  after ``install``); ``lru_cache``: RG-LRU state to continue from (h0);
  ``gate_mul``: fold the block's gating product into the output (fused kernel
  only).  ``out`` / ``last_h_out`` / ``conv_out`` / ``conv_cache_out``: optional
  caller-provided buffers.
  """
  if can_fuse_decode(conv, lru, x, conv_cache):
    # decode step: conv step + gate GEMVs + gates + h = a*h0 + x~ in ONE launch
    bsz = x.shape[0]
    if segment_pos.shape != (bsz, 1):
      segment_pos = segment_pos[None, :]
    assert segment_pos.shape == (bsz, 1)              # layers.py:344
    y, conv_state, h = _abi.recurrent_decode_step(
        x, conv.w, conv.b, conv_cache, lru.input_gate.w, lru.a_gate.w, lru.input_gate.b,
        lru.a_gate.b, lru.a_param, segment_pos, h0=lru_cache, gate_mul=gate_mul,
        return_cache=return_cache, arith_mode=layers.get_arith_mode())
    return y, conv_state, h
  if can_fuse_conv(conv, lru, x, conv_cache):
    # prefill: convolution + gate GEMMs + gate math + scan in ONE launch
    bsz, steps, _ = x.shape
    if segment_pos.shape != (bsz, steps):
      segment_pos = segment_pos[None, :]
    assert segment_pos.shape == (bsz, steps)            # layers.py:344
    if layers._empty_batch(x):
      xc, conv_state = layers._empty_conv1d(x, 4, None, return_cache, conv_out, conv_cache_out)
      y, h = layers._empty_rglru(xc, lru.width, return_cache, out, last_h_out)
      return y, conv_state, h
    y, conv_state, h = _abi.recurrent_prefill_fwd(
        x, conv.w, conv.b, layers.packed_gate_weight(lru), lru.input_gate.b, lru.a_gate.b, lru.a_param,
        segment_pos, lru.num_heads, h0=lru_cache, return_cache=return_cache,
        mask_mode=getattr(conv, "mask_mode", _abi.MASK_FORK), arith_mode=layers.get_arith_mode(),
        out=out, conv_cache_out=conv_cache_out, last_h_out=last_h_out, gate_mul=gate_mul)
    return y, conv_state, h
  xc, conv_state = (conv.forward_into(x, segment_pos, conv_cache, return_cache, out=conv_out,
                                      cache_out=conv_cache_out)
                    if hasattr(conv, "forward_into") else conv(x, segment_pos, conv_cache, return_cache))
  if gate_mul is not None and not layers.uses_fused_kernel(lru, xc):
    y, h = lru(xc, segment_pos, lru_cache, return_cache)
    return y * gate_mul, conv_state, h
  if hasattr(lru, "forward_into"):
    y, h = lru.forward_into(xc, segment_pos, lru_cache, return_cache, out=out,
                            last_h_out=last_h_out, gate_mul=gate_mul)
  else:   # a reference module patched by install()
    y, h = lru(xc, segment_pos, lru_cache, return_cache) if gate_mul is None else \
        type(lru).forward(lru, xc, segment_pos, lru_cache, return_cache, gate_mul=gate_mul)
  return y, conv_state, h


class GraphedHotPath:
  """``recurrent_hot_path`` for ONE fixed shape as a replayable CUDA graph (prefill).

  A small-batch prefill (the reference sampler's shape is B = 1, ``sampler.py:217-223``) is bound by
  the host: ~50 us of Python and launches per block step against 20-45 us of kernels.  This class
  captures the step once -- prologue + ONE fused launch (or the two / three kernels of the other
  routes) -- and replays it: no Python between the kernels, no allocator, no host synchronisation.

  Inputs are written into static buffers (``self.x`` ``[B,T,E]``, ``self.segment_pos`` ``[B,T]`` int32,
  ``self.h0`` ``[B,E]`` fp32 if ``with_h0``) -- either by the caller directly (a producer kernel that
  writes ``self.x`` saves a copy) or by ``__call__``, which copies them on the current stream;
  results live in ``self.y``, ``self.conv_state``, ``self.last_h`` until the next replay.
  """

  def __init__(self, conv, lru, batch: int, steps: int, with_h0: bool = False, warmup: int = 2):
    dev, dtype, width = conv.w.device, conv.w.dtype, lru.width
    self.conv, self.lru = conv, lru
    self.x = torch.zeros((batch, steps, width), device=dev, dtype=dtype)
    self.segment_pos = torch.arange(steps, device=dev, dtype=torch.int32)[None].repeat(batch, 1).contiguous()
    self.h0 = torch.zeros((batch, width), device=dev, dtype=torch.float32) if with_h0 else None
    self.y = torch.empty_like(self.x)
    self.conv_out = torch.empty_like(self.x)            # only touched by the two-kernel routes
    self.conv_state = torch.empty((batch, conv.w.shape[0] - 1, width), device=dev, dtype=dtype)
    self.last_h = torch.empty((batch, width), device=dev, dtype=torch.float32)
    self.device = dev
    # warm up ON the capture stream: the kernels' scratch buffers and the packed gate weights are
    # cached per stream / per weight version and must exist before the capture
    self._side = torch.cuda.Stream(dev)
    self._side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(self._side):
      for _ in range(max(1, warmup)):
        self._eager()
    self._side.synchronize()
    self._graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(self._graph, stream=self._side):
      self._eager()
    torch.cuda.current_stream(dev).wait_stream(self._side)

  @torch.no_grad()
  def _eager(self):
    _recurrent_hot_path_inference(self.conv, self.lru, self.x, self.segment_pos, None, self.h0, True, None,
                                  self.y, self.last_h, self.conv_out, self.conv_state)

  @torch.no_grad()
  def replay(self):
    """Runs the captured step on the contents of the static input buffers."""
    self._graph.replay()
    return self.y, self.conv_state, self.last_h

  @torch.no_grad()
  def __call__(self, x, segment_pos, h0=None):
    self.x.copy_(x)
    self.segment_pos.copy_(segment_pos if segment_pos.dim() == 2 else segment_pos[None].expand_as(self.segment_pos))
    if self.h0 is not None:
      if h0 is None:
        self.h0.zero_()
      else:
        self.h0.copy_(h0)
    else:
      assert h0 is None, "build with with_h0=True to continue from a state"
    return self.replay()
