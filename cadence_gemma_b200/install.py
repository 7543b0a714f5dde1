"""Drops the CUDA path into an imported copy of the reference.

``install(ref_layers, ref_modules)`` rebinds, inside the *reference's own*
modules, exactly the three entry points of the hot path:

  * ``layers.rnn_scan``        (looked up as a module global at call time,
                                reference layers.py:366)
  * ``layers.RGLRU.forward``   (reference layers.py:322-375)
  * ``layers.Conv1D.forward``  (reference layers.py:458-546)

and, given ``ref_modules``, ``modules.RecurrentBlock.forward`` (reference
modules.py:613-660), whose ``conv_1d -> rg_lru`` pair then runs as ONE fused
kernel (``cg_recurrent_prefill_fwd``) -- so an unmodified ``RecurrentBlock`` / ``ResidualBlock`` / ``Griffin`` built
from the reference classes (``examples/cadence_sampler.py`` included) runs the
sm_100a kernels with its own parameters.  ``uninstall`` restores the originals.
"""
from __future__ import annotations

import weakref

import torch

from cadence_gemma_b200 import _abi, layers as cg_layers

_saved = {}


_wcat_cache = weakref.WeakKeyDictionary()


def _fused_gate_gemm(lru, x):
  """One cuBLAS GEMM for both gates of a *reference* RGLRU module -> [N, H, 2*bw]."""
  wx, wa = lru.input_gate.w, lru.a_gate.w
  key = (wx.data_ptr(), wa.data_ptr(), wx._version, wa._version, wx.dtype, wx.device)
  hit = _wcat_cache.get(lru)
  if hit is None or hit[0] != key:
    hit = (key, torch.cat([wx.detach(), wa.detach()], dim=2).contiguous())
    _wcat_cache[lru] = hit
  heads, bw = lru.input_gate.num_blocks, lru.input_gate.block_width
  x2 = x.reshape(-1, heads, bw)
  out = torch.empty((x2.shape[0], heads, 2 * bw), dtype=x.dtype, device=x.device)
  torch.bmm(x2.transpose(0, 1), hit[1], out=out.transpose(0, 1))
  return out, bw


def _rglru_forward(self, x, segment_pos, cache=None, return_cache=True, gate_mul=None):
  bs, length, _ = x.shape
  if segment_pos.shape != (bs, length):
    segment_pos = segment_pos[None, :]
  assert segment_pos.shape == (bs, length)
  if cg_layers._empty_batch(x):
    return cg_layers._empty_rglru(x, self.width, return_cache)
  if cg_layers._wants_grad(x, cache, *self.parameters()):
    assert gate_mul is None
    if cg_layers._train_kernels and not (length == 1 and cache is None):
      # training kernels: the module's own gate GEMMs (autograd), then gate math +
      # scan forward / backward on our kernels
      y, last_h = cg_layers._RGLRUFn.apply(x, self.input_gate(x), self.a_gate(x), self.a_param,
                                           segment_pos == 0, cache)
      return (y, last_h) if return_cache else (y, None)
    # otherwise the reference's own forward builds the autograd graph; its scan is
    # the patched module-global rnn_scan = the differentiable kernel pair
    return _saved["rglru_forward"](self, x, segment_pos, cache, return_cache)
  with torch.no_grad():
    if cg_layers.uses_fused_kernel(self, x):
      # gate GEMMs + gate math + scan in ONE tcgen05 kernel (cg_rglru_fused_fwd)
      return _abi.rglru_fused_fwd(
          x, cg_layers.packed_gate_weight(self), self.input_gate.b, self.a_gate.b,
          self.a_param, segment_pos, self.num_heads, h0=cache, return_cache=return_cache,
          arith_mode=cg_layers.get_arith_mode(), gate_mul=gate_mul)
    assert gate_mul is None
    gates, bw = _fused_gate_gemm(self, x)
    return _abi.rglru_fwd(
        x, None, None, self.input_gate.b, self.a_gate.b, self.a_param, segment_pos,
        h0=cache, return_cache=return_cache, arith_mode=cg_layers.get_arith_mode(),
        gemm_fused=gates, block_width=bw)


def _conv1d_forward(self, x, segment_pos, cache=None, return_cache=True):
  mode = cg_layers.get_arith_mode() & _abi.ARITH_FP32
  if cg_layers._empty_batch(x):
    return cg_layers._empty_conv1d(x, self.temporal_width, cache, return_cache)
  if cg_layers._wants_grad(x, cache, self.w, self.b):
    if cache is not None:
      raise RuntimeError("Conv1D decode steps (cache given) are forward-only; call under torch.no_grad()")
    return cg_layers._Conv1DFn.apply(x, self.w, self.b, segment_pos, _abi.MASK_FORK, mode, return_cache)
  with torch.no_grad():
    if cache is not None:
      return _abi.conv1d_decode(x, self.w, self.b, cache,
                                return_cache=return_cache, arith_mode=mode)
    return _abi.conv1d_fwd(x, self.w, self.b, segment_pos,
                           return_cache=return_cache, arith_mode=mode)


def _recurrent_block_forward(self, x, segment_pos, cache=None, return_cache=True):
  """RecurrentBlock.forward (reference modules.py:613-660) with ``conv_1d -> rg_lru``
  (:638-649) routed through the hot-path entry point: ONE fused kernel for a bf16
  prefill at RecurrentGemma shapes (convolution inside the tcgen05 RG-LRU kernel),
  one launch for a decode step (gating product ``x * y``, :651, folded in there)."""
  from cadence_gemma_b200 import pipeline
  y = self.linear_y(x)
  h = self.linear_x(x)
  conv_cache = None if cache is None else cache.conv1d_state
  lru_cache = None if cache is None else cache.rg_lru_state
  grad = cg_layers._wants_grad(h, y, conv_cache, lru_cache, *self.conv_1d.parameters(),
                               *self.rg_lru.parameters())
  fold = not grad and (
      (cg_layers.fold_gate_enabled() and cg_layers.uses_fused_kernel(self.rg_lru, h)) or
      pipeline.can_fuse_decode(self.conv_1d, self.rg_lru, h, conv_cache))
  h, conv1d_state, rg_lru_state = pipeline.recurrent_hot_path(
      self.conv_1d, self.rg_lru, h, segment_pos, conv_cache=conv_cache, lru_cache=lru_cache,
      return_cache=return_cache, gate_mul=y if fold else None)
  if not fold:
    h = h * y
  out = self.linear_out(h)
  if not return_cache:
    return out, None
  return out, _saved["modules"].RecurrentBlockCache(
      conv1d_state=conv1d_state, rg_lru_state=rg_lru_state)


def install(ref_layers, ref_modules=None) -> None:
  """Patches the given reference modules in place (idempotent).  With
  ``ref_modules`` also ``RecurrentBlock.forward`` is rebound, so the gating
  product rides on the fused RG-LRU kernel."""
  if "rnn_scan" in _saved:
    return
  _abi.load()   # fail loudly now if the library was not built
  _saved["rnn_scan"] = ref_layers.rnn_scan
  _saved["rglru_forward"] = ref_layers.RGLRU.forward
  _saved["conv1d_forward"] = ref_layers.Conv1D.forward
  _saved["layers"] = ref_layers
  ref_layers.rnn_scan = cg_layers.rnn_scan
  ref_layers.RGLRU.forward = _rglru_forward
  ref_layers.Conv1D.forward = _conv1d_forward
  if ref_modules is not None and hasattr(ref_modules, "RecurrentBlock"):
    _saved["modules"] = ref_modules
    _saved["block_forward"] = ref_modules.RecurrentBlock.forward
    ref_modules.RecurrentBlock.forward = _recurrent_block_forward


def uninstall() -> None:
  if "rnn_scan" not in _saved:
    return
  ref_layers = _saved.pop("layers")
  ref_layers.rnn_scan = _saved.pop("rnn_scan")
  ref_layers.RGLRU.forward = _saved.pop("rglru_forward")
  ref_layers.Conv1D.forward = _saved.pop("conv1d_forward")
  if "modules" in _saved:
    _saved.pop("modules").RecurrentBlock.forward = _saved.pop("block_forward")
