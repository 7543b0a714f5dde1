"""Host-side mirror of the reference's recurrent layers, backed by sm_100a CUDA.

Same names, constructor arguments, parameter attributes / state-dict keys,
call signatures and return conventions as the reference
(``recurrentgemma/torch/layers.py``), so these classes drop into
``RecurrentBlock`` / ``ResidualBlock`` / ``Griffin``:

  * ``rnn_scan``             layers.py:146-199
  * ``BlockDiagonalLinear``  layers.py:81-142   (stays a cuBLAS GEMM)
  * ``RGLRU``                layers.py:241-386
  * ``Conv1D``               layers.py:389-676

The arithmetic runs in ``libcadence_b200.so`` through the C ABI
(``include/cadence_b200.h``).  There is no CPU fallback: CPU tensors or a
missing library raise.

Gradients: ``rnn_scan`` is differentiable --
its backward is the reverse-time scan kernel ``cg_rnn_scan_bwd`` -- and
``RGLRU.forward`` called with grad enabled runs gate math + scan forward and
backward on the training kernels (``_RGLRUFn``: ``cg_rglru_gates_fwd/bwd`` with the
clipped-sqrt gradient of layers.py:224-238, ``cg_rnn_scan_fwd/bwd``); the ``Conv1D`` prefill is differentiable too (``cg_conv1d_bwd``), so
``training/train.py`` can run on the kernels.  Decode steps are forward only.
"""
from __future__ import annotations

import math
import os
import weakref

import torch
from torch import nn

from cadence_gemma_b200 import _abi

# Default arithmetic of the kernels (see include/cadence_b200.h):
# reproduce every eager rounding point of the reference, fast transcendentals
# (99.8 % of the outputs bit-identical with the reference at every BASELINE config;
# _abi.ARITH_REFERENCE = exact transcendentals: 99.99 %, 1.7x the time).
DEFAULT_ARITH = _abi.ARITH_REFERENCE | _abi.ARITH_FAST
_arith_mode = DEFAULT_ARITH


def set_arith_mode(mode: int) -> int:
  """Sets the process-wide arithmetic mode of the shims; returns the old one."""
  global _arith_mode
  old, _arith_mode = _arith_mode, int(mode)
  return old


def get_arith_mode() -> int:
  return _arith_mode


# The fused tensor-core kernel (gate GEMMs + gate math + scan in one launch,
# cg_rglru_fused_fwd) is used whenever the shape allows it; CG_B200_FUSED=0 or
# set_fused(False) forces the cuBLAS-GEMM + scan-kernel pair instead.
_fused_enabled = os.environ.get("CG_B200_FUSED", "1") != "0"


def set_fused(enabled: bool) -> bool:
  """Enables / disables the fused tcgen05 RG-LRU path; returns the old setting."""
  global _fused_enabled
  old, _fused_enabled = _fused_enabled, bool(enabled)
  return old


def fused_enabled() -> bool:
  return _fused_enabled


# Folding RecurrentBlock's gating product `x * y` (reference modules.py:651) into
# the fused kernel's store (its `gate_mul` operand) is implemented
# and tested, but OFF by default: the operand arrives transposed to the kernel's
# thread layout (one 2-byte load per element in the latency-bound replay pass),
# which costs the kernel more (+50 us at config 2) than the separate elementwise
# multiply it saves (~40 us).  CG_B200_FOLD_GATE=1 / set_fold_gate(True) enable it.
_fold_gate = os.environ.get("CG_B200_FOLD_GATE", "0") != "0"


def set_fold_gate(enabled: bool) -> bool:
  global _fold_gate
  old, _fold_gate = _fold_gate, bool(enabled)
  return old


def fold_gate_enabled() -> bool:
  return _fold_gate


def _empty_batch(x) -> bool:
  """True for ``[0, T >= 1, E]`` activations (e.g. a rank whose batch shard holds no
  rows).  The reference runs its ops over zero rows and returns empty tensors
  (probe: ``rnn_scan`` / ``RGLRU`` / ``Conv1D`` with B = 0 give ``y [0,T,E]``,
  ``last_h [0,E]`` fp32, conv cache ``[0,W-1,E]``); there is nothing to launch.
  T = 0 stays an error, as in the reference (``x[:, 0]`` raises, layers.py:178)."""
  return x.shape[0] == 0 and x.shape[1] >= 1


def _empty_rglru(x, width, return_cache, out=None, last_h_out=None):
  y = out if out is not None else torch.empty_like(x)
  if not return_cache:
    return y, None
  return y, (last_h_out if last_h_out is not None else x.new_zeros((0, width), dtype=torch.float32))


def _empty_conv1d(x, temporal_width, cache, return_cache, out=None, cache_out=None):
  y = out if out is not None else torch.empty_like(x)
  if not return_cache:
    return y, None
  if cache_out is not None:
    return y, cache_out
  return y, x.new_zeros((0, temporal_width - 1, x.shape[2]),
                        dtype=x.dtype if cache is None else cache.dtype)   # decode keeps the cache's dtype (:542)


def _wants_grad(*tensors) -> bool:
  return torch.is_grad_enabled() and any(
      t is not None and t.requires_grad for t in tensors)


class _RnnScanFn(torch.autograd.Function):
  """``rnn_scan`` with the reverse-time scan kernel as its backward: the
  gradients autograd derives from the reference loop (layers.py:173, :187-199)."""

  @staticmethod
  def forward(ctx, x, a, reset, h0):
    y, last_h = _abi.rnn_scan_fwd(x, a, reset, h0, arith_mode=_arith_mode & _abi.ARITH_STRICT)
    ctx.has_h0 = h0 is not None
    ctx.save_for_backward(a, y, reset, h0 if h0 is not None else y.new_empty(0))
    return y, last_h

  @staticmethod
  def backward(ctx, gy, g_last):
    a, y, reset, h0 = ctx.saved_tensors
    if gy is None:
      gy = torch.zeros_like(y)
    want_h0 = ctx.has_h0 and ctx.needs_input_grad[3]
    dx, da, dh0 = _abi.rnn_scan_bwd(gy, g_last, a, y, reset, h0 if ctx.has_h0 else None,
                                    need_dh0=want_h0)
    return dx, da, None, dh0


class _RGLRUFn(torch.autograd.Function):
  """RG-LRU after the gate GEMMs, on the training kernels: forward = gate math
  (``cg_rglru_gates_fwd``) + scan (``cg_rnn_scan_fwd``); backward = reverse scan
  (``cg_rnn_scan_bwd``) + gate backward (``cg_rglru_gates_bwd``, incl. the
  clipped square-root gradient).  Saves the pre-activations, ``a`` and ``y``."""

  @staticmethod
  def forward(ctx, x, pre_x, pre_a, a_param, reset, h0):
    a, nx = _abi.rglru_gates_fwd(x, pre_x, pre_a, a_param, reset)
    y, last_h = _abi.rnn_scan_fwd(nx, a, reset, h0, arith_mode=0)
    ctx.has_h0 = h0 is not None
    ctx.save_for_backward(x, pre_x, pre_a, a_param, reset, a, y,
                          h0 if h0 is not None else y.new_empty(0))
    return y, last_h

  @staticmethod
  def backward(ctx, gy, g_last):
    x, pre_x, pre_a, a_param, reset, a, y, h0 = ctx.saved_tensors
    if gy is None:
      gy = torch.zeros_like(y)
    want_h0 = ctx.has_h0 and ctx.needs_input_grad[5]
    d_nx, d_a, dh0 = _abi.rnn_scan_bwd(gy, g_last, a, y, reset, h0 if ctx.has_h0 else None,
                                       need_dh0=want_h0)
    dx, dpx, dpa, dap = _abi.rglru_gates_bwd(x, pre_x, pre_a, a_param, reset, d_nx, d_a)
    return dx, dpx, dpa, dap, None, dh0


_train_kernels = os.environ.get("CG_B200_TRAIN_KERNELS", "1") != "0"


def set_train_kernels(enabled: bool) -> bool:
  """Training path of ``RGLRU.forward``: True (default) = gate math forward and
  backward on our kernels (``_RGLRUFn``); False = the reference's ATen op sequence
  under autograd around the differentiable scan.  Returns the previous setting."""
  global _train_kernels
  prev, _train_kernels = _train_kernels, bool(enabled)
  return prev


class SqrtBoundDerivative(torch.autograd.Function):
  """``sqrt`` whose gradient is bounded by ``_MAX_SQRT_GRADIENT`` (reference
  layers.py:224-238): d/dx sqrt(x) = 1 / sqrt(max(4 x, 1 / bound^2))."""

  @staticmethod
  def forward(ctx, x):
    ctx.save_for_backward(x)
    return torch.sqrt(x)

  @staticmethod
  def backward(ctx, grad_output):
    (x,) = ctx.saved_tensors
    return grad_output / torch.sqrt(torch.clip(4.0 * x, min=1.0 / (_MAX_SQRT_GRADIENT ** 2)))


_MAX_SQRT_GRADIENT = 1000.0


def rnn_scan(x, a, reset, h0, acc_dtype=torch.float32):
  """Linear recurrence h_t = a_t h_{t-1} + x_t (reference layers.py:146-199).

  Returns ``(y in x.dtype, h_last in fp32)``.  Differentiable in ``x``, ``a``
  and ``h0`` (backward = ``cg_rnn_scan_bwd``).
  """
  assert x.ndim == 3
  assert a.shape == x.shape[-a.ndim:]
  assert a.dtype == x.dtype
  assert acc_dtype == torch.float32, "only fp32 accumulation is implemented"
  assert h0 is None or h0.dtype == acc_dtype
  if a.shape != x.shape:
    a = a.expand_as(x)
  if _empty_batch(x):
    return x.clone(), x.new_zeros((0, x.shape[2]), dtype=acc_dtype)
  if _wants_grad(x, a, h0):
    if x.shape[1] == 1 and h0 is None:      # :177-178: the output IS x (views, no arithmetic)
      return x, x[:, 0].type(acc_dtype)
    return _RnnScanFn.apply(x, a, reset, h0)
  mode = _arith_mode & _abi.ARITH_STRICT
  return _abi.rnn_scan_fwd(x, a, reset, h0, arith_mode=mode)


_wpack_cache = weakref.WeakKeyDictionary()


def packed_gate_weight(lru) -> torch.Tensor:
  """Gate weights of an RGLRU-like module (ours or the reference's) as the
  shared-memory image the fused tcgen05 kernel reads
  (cg_rglru_pack_gate_weights); rebuilt only when a gate weight changes."""
  wx, wa = lru.input_gate.w, lru.a_gate.w
  key = (wx.data_ptr(), wa.data_ptr(), wx._version, wa._version, wx.dtype, wx.device)
  hit = _wpack_cache.get(lru)
  if hit is None or hit[0] != key:
    hit = (key, _abi.pack_gate_weights(wx, wa))
    _wpack_cache[lru] = hit
  return hit[1]


def uses_fused_kernel(lru, x: torch.Tensor) -> bool:
  """True if the RG-LRU of module ``lru`` on input ``x`` runs on the fused
  tensor-core kernel (bf16, head width 128 / 256, prefill, emulated modes)."""
  return (_fused_enabled and x.is_cuda and x.shape[1] > 1 and
          (_arith_mode & (_abi.ARITH_FP32 | _abi.ARITH_STRICT)) == 0 and
          _abi.fused_supported(lru.width, lru.num_heads, x.dtype))


class _BlockDiagonalLinearFn(torch.autograd.Function):
  """``x @ blockdiag(w) + b`` for the training path: strided-batched cuBLAS GEMMs
  that read and write the flat ``[..., H*bw]`` layout directly, forward and
  backward (autograd through ``einsum`` / ``bmm`` + ``transpose`` materialises a
  permuted copy of every activation-sized tensor)."""

  @staticmethod
  def forward(ctx, x, w, b):
    heads, bw, _ = w.shape
    x2 = x.reshape(-1, heads, bw)
    out = torch.empty_like(x2)
    torch.bmm(x2.transpose(0, 1), w, out=out.transpose(0, 1))
    out += b                                   # r(r(x @ w) + b), layers.py:139
    ctx.save_for_backward(x2, w)
    return out.view(x.shape)

  @staticmethod
  def backward(ctx, gy):
    x2, w = ctx.saved_tensors
    heads, bw, _ = w.shape
    gy2 = gy.reshape(-1, heads, bw)
    dx = dw = db = None
    if ctx.needs_input_grad[0]:
      dx2 = torch.empty_like(x2)
      torch.bmm(gy2.transpose(0, 1), w.transpose(1, 2), out=dx2.transpose(0, 1))
      dx = dx2.view(gy.shape)
    if ctx.needs_input_grad[1]:
      dw = torch.bmm(x2.transpose(0, 1).transpose(1, 2), gy2.transpose(0, 1))
    if ctx.needs_input_grad[2]:
      db = gy2.sum(0)
    return dx, dw, db


class BlockDiagonalLinear(nn.Module):
  """Block-diagonal linear layer (reference layers.py:81-142)."""

  def __init__(self, width, num_blocks, w_init_variance_scale=1.0, device=None,
               dtype=None):
    super().__init__()
    self.width = width
    self.num_blocks = num_blocks
    self.w_init_variance_scale = w_init_variance_scale
    self.block_width = width // num_blocks
    bw = self.block_width
    self.w = nn.Parameter(torch.empty((num_blocks, bw, bw), device=device, dtype=dtype))
    self.b = nn.Parameter(torch.empty((num_blocks, bw), device=device, dtype=dtype))
    self.reset_parameters()

  def reset_parameters(self) -> None:
    self.w_init_(self.w)
    nn.init.zeros_(self.b)

  def w_init_(self, w: torch.Tensor) -> None:
    nn.init.normal_(w, mean=0.0,
                    std=math.sqrt(self.w_init_variance_scale / self.block_width))

  def gemm(self, x: torch.Tensor) -> torch.Tensor:
    """``x @ blockdiag(w)`` WITHOUT bias, laid out like ``x`` ([..., H*bw]).

    One strided-batched cuBLAS GEMM writing straight into the flat layout (no
    permute / reshape copies); the bias is added inside the scan kernel.
    """
    heads, bw = self.num_blocks, self.block_width
    x2 = x.reshape(-1, heads, bw)
    out = torch.empty_like(x2)
    torch.bmm(x2.transpose(0, 1), self.w, out=out.transpose(0, 1))
    return out.view(x.shape)

  def forward(self, x: torch.Tensor) -> torch.Tensor:
    heads, bw = self.num_blocks, self.block_width
    if _wants_grad(x, self.w, self.b):          # training path
      return _BlockDiagonalLinearFn.apply(x, self.w, self.b)
    y = self.gemm(x).view(*x.shape[:-1], heads, bw) + self.b
    return y.view(x.shape)


class RGLRU(nn.Module):
  """Real-Gated Linear Recurrent Unit (reference layers.py:241-386)."""

  def __init__(self, width, num_heads, w_init_variance_scale=1.0, device=None,
               dtype=None):
    super().__init__()
    self.width = width
    self.num_heads = num_heads
    self.w_init_variance_scale = w_init_variance_scale
    self.a_param = nn.Parameter(torch.empty((width,), device=device, dtype=dtype))
    self.input_gate = BlockDiagonalLinear(width, num_heads, w_init_variance_scale,
                                          device=device, dtype=dtype)
    self.a_gate = BlockDiagonalLinear(width, num_heads, w_init_variance_scale,
                                      device=device, dtype=dtype)
    self.reset_parameters()

  def reset_parameters(self) -> None:
    self.input_gate.reset_parameters()
    self.a_gate.reset_parameters()
    self.a_param_init(self.a_param)

  def a_param_init(self, w: torch.Tensor) -> torch.Tensor:
    """`a` uniform on the ring [0.9, 0.999], softplus-inverse (layers.py:202-221)."""
    lo, hi, eps = 0.9, 0.999, 1e-8
    with torch.no_grad():
      w.uniform_(lo * lo + eps, hi * hi + eps)
      w.log_().mul_(0.5)
      return w.neg_().exp_().sub_(1.0).log_()

  def _fused_gate_weight(self) -> torch.Tensor:
    """[H, bw, 2*bw]: input-gate and a-gate weights side by side, so ONE
    strided-batched cuBLAS GEMM reads the activations once for both gates.
    Rebuilt only when a gate weight changes (in-place update or re-assignment)."""
    wx, wa = self.input_gate.w, self.a_gate.w
    key = (wx.data_ptr(), wa.data_ptr(), wx._version, wa._version, wx.dtype, wx.device)
    if getattr(self, "_wcat_key", None) != key:
      self._wcat = torch.cat([wx.detach(), wa.detach()], dim=2).contiguous()
      self._wcat_key = key
    return self._wcat

  def uses_fused_kernel(self, x: torch.Tensor) -> bool:
    """True if ``forward(x, ...)`` runs on the fused tensor-core kernel."""
    return uses_fused_kernel(self, x)

  def gate_gemm(self, x: torch.Tensor) -> torch.Tensor:
    """Both gate GEMMs (no bias) in one call -> [B*T, H, 2*bw]."""
    heads, bw = self.num_heads, self.width // self.num_heads
    x2 = x.reshape(-1, heads, bw)
    out = torch.empty((x2.shape[0], heads, 2 * bw), dtype=x.dtype, device=x.device)
    torch.bmm(x2.transpose(0, 1), self._fused_gate_weight(), out=out.transpose(0, 1))
    return out

  def forward(self, x, segment_pos, cache=None, return_cache=True):
    """Returns ``(y in x.dtype, last_h fp32 | None)``."""
    return self.forward_into(x, segment_pos, cache, return_cache)

  def forward_into(self, x, segment_pos, cache=None, return_cache=True, out=None,
                   last_h_out=None, gate_mul=None):
    """``forward`` writing into caller-provided ``out`` / ``last_h_out`` buffers
    (no allocation on the hot path; used by ``hostio.HostPrefill``).

    ``gate_mul`` (fused kernel only): return ``y * gate_mul`` rounded to bf16,
    the gating product of ``RecurrentBlock.forward`` (modules.py:651)."""
    bs, length, _ = x.shape
    if segment_pos.shape != (bs, length):
      segment_pos = segment_pos[None, :]
    assert segment_pos.shape == (bs, length)      # layers.py:344
    if _empty_batch(x):
      return _empty_rglru(x, self.width, return_cache, out, last_h_out)
    if _wants_grad(x, cache, *self.parameters()):
      assert out is None and last_h_out is None and gate_mul is None
      return self._forward_autograd(x, segment_pos, cache, return_cache)
    with torch.no_grad():
      if self.uses_fused_kernel(x):
        return _abi.rglru_fused_fwd(
            x, packed_gate_weight(self), self.input_gate.b, self.a_gate.b,
            self.a_param, segment_pos, self.num_heads, h0=cache,
            return_cache=return_cache, arith_mode=_arith_mode, out=out,
            last_h_out=last_h_out, gate_mul=gate_mul)
      assert gate_mul is None, "the gating product is folded in only by the fused kernel"
      y, last_h = _abi.rglru_fwd(
          x, None, None, self.input_gate.b, self.a_gate.b, self.a_param,
          segment_pos, h0=cache, return_cache=return_cache,
          arith_mode=_arith_mode, gemm_fused=self.gate_gemm(x),
          block_width=self.width // self.num_heads, out=out, last_h_out=last_h_out)
    return y, last_h

  def _forward_autograd(self, x, segment_pos, cache, return_cache):
    """Training path (grad enabled and x / cache / a parameter requires grad):
    the reference's op sequence (layers.py:345-371) as ATen ops, so autograd
    sees the same graph, around the differentiable scan kernels."""
    reset = segment_pos == 0
    if _train_kernels and not (x.shape[1] == 1 and cache is None):
      # gate GEMMs by cuBLAS under autograd; everything after them on our kernels
      y, last_h = _RGLRUFn.apply(x, self.input_gate(x), self.a_gate(x), self.a_param, reset, cache)
      return (y, last_h) if return_cache else (y, None)
    gate_x = torch.sigmoid(self.input_gate(x))
    gate_a = torch.sigmoid(self.a_gate(x))
    log_a = -8.0 * gate_a * nn.functional.softplus(self.a_param)
    a = torch.exp(log_a)
    a_square = torch.exp(2 * log_a)
    gated_x = x * gate_x
    multiplier = SqrtBoundDerivative.apply(1 - a_square)
    multiplier = reset[..., None] + ~reset[..., None] * multiplier
    normalized_x = gated_x * multiplier.type(x.dtype)
    y, last_h = rnn_scan(normalized_x, a, reset, cache)
    return (y, last_h) if return_cache else (y, None)

  @classmethod
  def init_cache(cls, batch_size, width, device=None):
    return torch.zeros((batch_size, width), dtype=torch.float32, device=device)


class _Conv1DFn(torch.autograd.Function):
  """Conv1D prefill with ``cg_conv1d_bwd`` as its backward.  The returned cache
  (a copy of the last input rows) is not differentiable."""

  @staticmethod
  def forward(ctx, x, w, b, segment_pos, mask_mode, arith_mode, return_cache):
    y, cache = _abi.conv1d_fwd(x, w, b, segment_pos, return_cache=return_cache,
                               mask_mode=mask_mode, arith_mode=arith_mode)
    ctx.save_for_backward(x, w, segment_pos)
    ctx.mask_mode = mask_mode
    if cache is None:
      return y, None
    ctx.mark_non_differentiable(cache)
    return y, cache

  @staticmethod
  def backward(ctx, gy, _g_cache):
    x, w, segment_pos = ctx.saved_tensors
    dx, dw, db = _abi.conv1d_bwd(gy, x, w, segment_pos, mask_mode=ctx.mask_mode)
    return dx, dw, db, None, None, None, None


class Conv1D(nn.Module):
  """Depth-wise causal temporal convolution (reference layers.py:389-676)."""

  def __init__(self, width, temporal_width, w_init_variance_scale=0.01,
               device=None, dtype=None):
    super().__init__()
    self.width = width
    self.temporal_width = temporal_width
    self.w_init_variance_scale = w_init_variance_scale
    self.w = nn.Parameter(torch.empty((temporal_width, width), device=device, dtype=dtype))
    self.b = nn.Parameter(torch.empty((width,), device=device, dtype=dtype))
    # CG_MASK_FORK reproduces the fork's document mask (layers.py:629-632);
    # set to _abi.MASK_UPSTREAM for the upstream / JAX behaviour.
    self.mask_mode = _abi.MASK_FORK
    self.reset_parameters()

  def reset_parameters(self) -> None:
    self.w_init_(self.w)
    nn.init.zeros_(self.b)

  def w_init_(self, w: torch.Tensor) -> None:
    nn.init.normal_(w, mean=0.0,
                    std=math.sqrt(self.w_init_variance_scale / self.temporal_width))

  def forward(self, x, segment_pos, cache=None, return_cache=True):
    """Returns ``(y, new_cache | None)``; decode when ``cache`` is given."""
    return self.forward_into(x, segment_pos, cache, return_cache)

  def forward_into(self, x, segment_pos, cache=None, return_cache=True, out=None,
                   cache_out=None):
    """``forward`` (prefill) writing into caller-provided buffers."""
    mode = _arith_mode & (_abi.ARITH_FP32)
    if _empty_batch(x):
      return _empty_conv1d(x, self.temporal_width, cache, return_cache, out, cache_out)
    if _wants_grad(x, cache, self.w, self.b):
      # training path: differentiable prefill (temporal width 4); decoding with a
      # cache is an inference-only operation
      if cache is not None:
        raise RuntimeError("Conv1D decode steps (cache given) are forward-only; call under torch.no_grad()")
      assert out is None and cache_out is None
      return _Conv1DFn.apply(x, self.w, self.b, segment_pos, self.mask_mode, mode, return_cache)
    with torch.no_grad():
      if cache is not None:
        return _abi.conv1d_decode(x, self.w, self.b, cache,
                                  return_cache=return_cache, arith_mode=mode)
      return _abi.conv1d_fwd(x, self.w, self.b, segment_pos,
                             return_cache=return_cache, mask_mode=self.mask_mode,
                             arith_mode=mode, out=out, cache_out=cache_out)

  @classmethod
  def init_cache(cls, *, batch_size, width, dtype, conv1d_temporal_width=4,
                 device=None):
    return torch.zeros((batch_size, conv1d_temporal_width - 1, width),
                       dtype=dtype, device=device)
