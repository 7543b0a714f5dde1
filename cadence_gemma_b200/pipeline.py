"""Conv1D -> RG-LRU as ONE overlapped pipeline (prefill).

``recurrent_hot_path(conv, lru, x, segment_pos)`` computes what
``RecurrentBlock.forward`` computes between its projections (reference
``recurrentgemma/torch/modules.py:638-649``):

    x, conv1d_state = conv_1d(x, segment_pos)
    x, rg_lru_state = rg_lru(x, segment_pos)

By default the two kernels run one after the other on the caller's stream.

EXPERIMENTAL (``set_overlap(True)`` / ``CG_B200_OVERLAP=1``): the temporal
convolution is HBM-bound while the fused tensor-core RG-LRU kernel is bound by
the issue / MUFU rate of its gate math and leaves most of the HBM bandwidth idle,
so the two can run AT THE SAME TIME: the convolution as a few persistent producer
blocks on a side stream (``cg_conv1d_stream_fwd``, time-major order, a counter
per 64-step group and batch row), the RG-LRU kernel on the caller's stream with
its TMA producer waiting on those counters tile by tile (``conv_flags`` of
``cg_rglru_fused_fwd``).  The protocol is correct (tested bit for bit against the
sequential pair, with the intermediate poisoned) but on B200 it is SLOWER
(287 us vs 137 us at config 2): the fused kernel's 576 threads x 96 registers
leave room for one 128-thread convolution block per SM, far too few loads in
flight for an HBM-bound producer.  Kept switched off; see DESIGN.md section 9.

No CPU path.
"""
from __future__ import annotations

import os

import torch

from cadence_gemma_b200 import _abi, layers

_overlap = os.environ.get("CG_B200_OVERLAP", "0") != "0"
# One-launch decode step (cg_recurrent_decode_step).  It more than halves the cost
# of an EAGER decode step (30 us vs 76 us per block at B = 32: three launches and
# their host overhead become one); inside a CUDA graph, where launch overhead is
# gone, the three small kernels are ~5 us faster at B = 32 -- switch it off there.
_fused_decode = os.environ.get("CG_B200_FUSED_DECODE", "1") != "0"
_side_streams: dict = {}
_flag_buffers: dict = {}


def _side_stream(device) -> torch.cuda.Stream:
  s = _side_streams.get(device)
  if s is None:
    s = torch.cuda.Stream(device)
    _side_streams[device] = s
  return s


def set_overlap(enabled: bool) -> bool:
  """Switches the experimental producer / consumer overlap; returns the old setting."""
  global _overlap
  old, _overlap = _overlap, bool(enabled)
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


def can_overlap(conv, lru, x, conv_cache=None) -> bool:
  """True if ``recurrent_hot_path`` runs the convolution under the RG-LRU kernel."""
  return (_overlap and layers.fused_enabled() and conv_cache is None and x.is_cuda and
          conv.w.shape[0] == 4 and x.shape[-1] % 64 == 0 and
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
  if not can_overlap(conv, lru, x, conv_cache):
    xc, conv_state = conv(x, segment_pos, conv_cache, return_cache)
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

  bsz, steps, width = x.shape
  if segment_pos.shape != (bsz, steps):
    segment_pos = segment_pos[None, :]
  assert segment_pos.shape == (bsz, steps)            # layers.py:344
  dev = x.device
  cur = torch.cuda.current_stream(dev)
  side = _side_stream(dev)
  x = x.contiguous()
  key = (dev, cur.cuda_stream, bsz, steps)
  flags = _flag_buffers.get(key)
  if flags is None:
    n = _abi.load().cg_conv1d_stream_flags_bytes(bsz, steps) // 4
    flags = torch.zeros(n, dtype=torch.int32, device=dev)
    if len(_flag_buffers) > 64:
      _flag_buffers.clear()
    _flag_buffers[key] = flags
  xc = torch.empty_like(x) if conv_out is None else conv_out
  conv_state = None
  if return_cache:
    conv_state = (torch.empty((bsz, 3, width), dtype=x.dtype, device=dev)
                  if conv_cache_out is None else conv_cache_out)
  wpack = layers.packed_gate_weight(lru)               # (re)packs on the caller's stream if stale
  flags.zero_()                                        # consumer's stream, before the fork
  side.wait_stream(cur)
  with torch.cuda.stream(side):                        # producer FIRST (it never waits for the consumer)
    _abi.conv1d_stream_fwd(x, conv.w, conv.b, segment_pos, flags, out=xc, cache_out=conv_state,
                           mask_mode=getattr(conv, "mask_mode", _abi.MASK_FORK))
  y, h = _abi.rglru_fused_fwd(xc, wpack, lru.input_gate.b, lru.a_gate.b, lru.a_param, segment_pos,
                              lru.num_heads, h0=lru_cache, return_cache=return_cache,
                              arith_mode=layers.get_arith_mode(), out=out, last_h_out=last_h_out,
                              gate_mul=gate_mul, conv_flags=flags)
  cur.wait_stream(side)                                # join: conv_state / xc are complete on `cur`
  return y, conv_state, h
