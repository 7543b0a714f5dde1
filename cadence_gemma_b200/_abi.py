"""ctypes binding of ``libcadence_b200.so`` (C ABI in ``include/cadence_b200.h``).

There is no CPU or eager-torch fallback: if the library is missing or a call
fails, an exception is raised.  PyTorch is used only for device memory and the
current CUDA stream.
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CG_B200_LIB") or os.path.join(_HERE, "csrc", "libcadence_b200.so")

DTYPE_F32, DTYPE_BF16 = 0, 1
ARITH_REFERENCE, ARITH_FP32, ARITH_FAST, ARITH_STRICT = 0, 1, 2, 4
MASK_FORK, MASK_UPSTREAM = 0, 1

# Every symbol include/cadence_b200.h declares: (name, restype, argtypes)
_vp, _i, _ll, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_size_t
SYMBOLS = {
    "cg_abi_version": (_i, []),
    "cg_status_string": (ctypes.c_char_p, [_i]),
    "cg_scan_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "cg_conv1d_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _ll, _vp, _vp, _i, _i, _i, _i,
                           _i, _i, _i, _vp]),
    "cg_conv1d_decode": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i,
                              _i, _vp]),
    "cg_rglru_fwd": (_i, [_vp, _vp, _vp, _ll, _i, _vp, _vp, _vp, _vp, _i, _ll, _vp,
                          _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _vp]),
    "cg_rnn_scan_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i,
                             _i, _i, _vp]),
    "cg_conv1d_bwd_workspace_bytes": (_sz, [_i, _i, _i]),
    "cg_conv1d_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _ll, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i,
                           _i, _i, _vp]),
    "cg_rglru_gates_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "cg_rglru_gates_bwd_workspace_bytes": (_sz, [_i, _i, _i]),
    "cg_rglru_gates_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz,
                                _i, _i, _i, _i, _vp]),
    "cg_rnn_scan_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i,
                             _i, _vp]),
    "cg_rglru_fused_supported": (_i, [_i, _i, _i]),
    "cg_rglru_gate_pack_bytes": (_sz, [_i, _i]),
    "cg_rglru_pack_gate_weights": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "cg_rglru_fused_workspace_bytes": (_sz, [_i, _i, _i]),
    "cg_rglru_fused_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _ll, _vp, _vp, _vp,
                                _vp, _sz, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "cg_recurrent_prefill_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _ll, _vp, _vp, _vp, _vp,
                                      _vp, _vp, _sz, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "cg_recurrent_decode_supported": (_i, [_i, _i, _i, _i]),
    "cg_recurrent_decode_step": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _ll,
                                      _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
}
# include/cadence_b200_experimental.h (development taps, not part of the drop-in surface)
EXPERIMENTAL_SYMBOLS = {
    "cg_recurrent_prefill_debug": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _ll, _vp, _vp, _vp, _vp,
                                        _vp, _sz, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
}
ABI_VERSION = 2

_lib = None
_lock = threading.Lock()
launch_count = 0   # kernels launched through this binding (bench bookkeeping)


class CadenceAbiError(RuntimeError):
  pass


def load():
  """Loads the shared library (once).  Raises if it has not been built."""
  global _lib
  if _lib is not None:
    return _lib
  with _lock:
    if _lib is None:
      if not os.path.exists(LIB_PATH):
        raise CadenceAbiError(
            f"{LIB_PATH} is missing: run `python -m cadence_gemma_b200.build` "
            "(there is no CPU / eager fallback for the recurrent hot path)")
      lib = ctypes.CDLL(LIB_PATH)
      for name, (restype, argtypes) in {**SYMBOLS, **EXPERIMENTAL_SYMBOLS}.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = restype, argtypes
      if lib.cg_abi_version() != ABI_VERSION:
        raise CadenceAbiError(
            f"libcadence_b200.so ABI version {lib.cg_abi_version()} != {ABI_VERSION}: rebuild "
            "(python -m cadence_gemma_b200.build --force)")
      _lib = lib
  return _lib


def _check(rc: int, what: str):
  if rc != 0:
    msg = load().cg_status_string(rc).decode()
    if rc < 0:
      # argument errors mirror the reference's `assert`s
      raise AssertionError(f"{what}: {msg} (status {rc})")
    raise CadenceAbiError(f"{what}: CUDA error {rc}: {msg}")


def dtype_code(dtype: torch.dtype) -> int:
  if dtype == torch.float32:
    return DTYPE_F32
  if dtype == torch.bfloat16:
    return DTYPE_BF16
  raise TypeError(f"cadence_gemma_b200 supports float32/bfloat16, got {dtype}")


def _ptr(t):
  return None if t is None else t.data_ptr()


def _stream(t: torch.Tensor):
  return torch.cuda.current_stream(t.device).cuda_stream


class _NoSwitch:
  """Stand-in for ``torch.cuda.device`` when the tensor already lives on the current
  device: the device-guard context costs several microseconds per call, which is
  what a small-batch prefill (host-bound: ~15 us of Python per kernel) is made of."""
  __slots__ = ()

  def __enter__(self):
    return None

  def __exit__(self, *exc):
    return False


_NO_SWITCH = _NoSwitch()


def _on_device(device):
  if device.index is None or device.index == torch.cuda.current_device():
    return _NO_SWITCH
  return torch.cuda.device(device)


def _require_cuda(*tensors):
  for t in tensors:
    if t is not None and not t.is_cuda:
      raise CadenceAbiError(
          "the recurrent hot path runs on CUDA only (no CPU fallback); got a "
          f"{t.device} tensor")


def _seg_args(segment_pos: torch.Tensor, batch: int, steps: int):
  """(tensor kept alive, is_i64, batch stride) for a [B,T] / [1,T] / [T] tensor."""
  seg = segment_pos
  if seg.dtype not in (torch.int32, torch.int64):
    seg = seg.to(torch.int64)
  if seg.ndim == 1:
    seg = seg[None, :]
  assert seg.ndim == 2 and seg.shape[1] == steps, (tuple(seg.shape), steps)
  assert seg.shape[0] in (1, batch), (tuple(seg.shape), batch)
  seg = seg.contiguous()
  stride = 0 if seg.shape[0] == 1 else steps
  return seg, int(seg.dtype == torch.int64), stride


# ---------------------------------------------------------------------------
# Scratch buffers: ONE grow-only buffer per (device, stream, kind).  The kernels tag
# every exchange word with a per-launch epoch, so a buffer can be reused across
# shapes; it is never freed or replaced by a smaller one while the process lives,
# and a buffer that had to grow is kept alive next to its successor (a captured
# CUDA graph may still hold its address -- ADVICE r1).
_workspaces: dict = {}
_retired_workspaces: list = []


def _grow_only_workspace(kind: str, device, nbytes: int) -> torch.Tensor:
  key = (kind, device, torch.cuda.current_stream(device).cuda_stream)
  ws = _workspaces.get(key)
  if ws is None or ws.numel() < nbytes:
    if ws is not None:
      _retired_workspaces.append(ws)
    ws = torch.zeros(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)   # zero-filled ONCE (header contract)
    _workspaces[key] = ws
  return ws


def scan_workspace(device, batch: int, steps: int, width: int, dtype) -> torch.Tensor:
  """Scratch of the scan kernels for the current stream (grow-only, zero-filled once)."""
  nbytes = load().cg_scan_workspace_bytes(batch, steps, width, dtype_code(dtype))
  return _grow_only_workspace("scan", device, nbytes)


def conv1d_fwd(x, w, b, segment_pos, return_cache=True, mask_mode=MASK_FORK,
               arith_mode=ARITH_REFERENCE, out=None, cache_out=None):
  global launch_count
  _require_cuda(x, w, b, segment_pos)
  bsz, steps, width = x.shape
  tw = w.shape[0]
  x, w, b = x.contiguous(), w.contiguous(), b.contiguous()
  seg, is64, stride = _seg_args(segment_pos, bsz, steps)
  y = torch.empty_like(x) if out is None else out
  cache = None
  if return_cache:
    cache = (torch.empty((bsz, tw - 1, width), dtype=x.dtype, device=x.device)
             if cache_out is None else cache_out)
  assert y.shape == x.shape and y.dtype == x.dtype and y.is_contiguous()
  with _on_device(x.device):
    rc = load().cg_conv1d_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(),
                              seg.data_ptr(), is64, stride, y.data_ptr(),
                              _ptr(cache), bsz, steps, width, tw,
                              dtype_code(x.dtype), mask_mode, arith_mode,
                              _stream(x))
  _check(rc, "cg_conv1d_fwd")
  launch_count += 1
  return y, cache


def conv1d_bwd(gy, x, w, segment_pos, mask_mode=MASK_FORK):
  """Backward of the Conv1D prefill (cg_conv1d_bwd): returns ``(dx, dw, db)``."""
  global launch_count
  _require_cuda(gy, x, w, segment_pos)
  bsz, steps, width = x.shape
  assert gy.shape == x.shape and gy.dtype == x.dtype and w.dtype == x.dtype
  gy, x, w = gy.contiguous(), x.contiguous(), w.contiguous()
  seg, is64, stride = _seg_args(segment_pos, bsz, steps)
  dx = torch.empty_like(x)
  dw = torch.empty_like(w)
  db = torch.empty((width,), dtype=x.dtype, device=x.device)
  nbytes = load().cg_conv1d_bwd_workspace_bytes(bsz, steps, width)
  ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
  with _on_device(x.device):
    rc = load().cg_conv1d_bwd(gy.data_ptr(), x.data_ptr(), w.data_ptr(), seg.data_ptr(), is64, stride,
                              dx.data_ptr(), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), ws.numel(),
                              bsz, steps, width, w.shape[0], dtype_code(x.dtype), mask_mode, _stream(x))
  _check(rc, "cg_conv1d_bwd")
  launch_count += 2
  return dx, dw, db


def conv1d_decode(x, w, b, cache, return_cache=True, arith_mode=ARITH_REFERENCE, cache_out=None):
  """``cache_out`` may be ``cache`` itself (in-place roll; the kernel reads a thread's rows
  before it writes them -- include/cadence_b200.h)."""
  global launch_count
  _require_cuda(x, w, b, cache)
  bsz, steps, width = x.shape
  tw = w.shape[0]
  assert steps == 1, "layers.py:566: decode takes exactly one token"
  assert cache.shape == (bsz, tw - 1, width), "layers.py:565"
  x, w, b, cache = x.contiguous(), w.contiguous(), b.contiguous(), cache.contiguous()
  y = torch.empty_like(x)
  new_cache = (torch.empty_like(cache) if cache_out is None else cache_out) if return_cache else None
  with _on_device(x.device):
    rc = load().cg_conv1d_decode(x.data_ptr(), w.data_ptr(), b.data_ptr(),
                                 cache.data_ptr(), dtype_code(cache.dtype),
                                 y.data_ptr(), _ptr(new_cache), bsz, width, tw,
                                 dtype_code(x.dtype), arith_mode, _stream(x))
  _check(rc, "cg_conv1d_decode")
  launch_count += 1
  return y, new_cache


def rglru_fwd(x, gemm_x, gemm_a, bias_x, bias_a, a_param, segment_pos, h0=None,
              return_cache=True, arith_mode=ARITH_REFERENCE, out=None,
              gemm_fused=None, block_width=0, last_h_out=None):
  """Gate math + scan.

  Either ``gemm_x`` / ``gemm_a`` ([B,T,E], possibly row-strided views with unit
  inner stride), or ``gemm_fused``: the [B*T, H, 2*bw] output of ONE fused
  block-diagonal GEMM (per head: input-gate columns, then a-gate columns) with
  ``block_width = bw``.
  """
  global launch_count
  _require_cuda(x, gemm_x, gemm_a, gemm_fused, a_param, segment_pos, h0)
  bsz, steps, width = x.shape
  x = x.contiguous()
  assert a_param.dtype == x.dtype, "a_param must have the activation dtype"
  assert h0 is None or h0.dtype == torch.float32, "layers.py:170"
  if gemm_fused is not None:
    assert block_width > 0 and width % block_width == 0
    assert gemm_fused.is_contiguous() and gemm_fused.dtype == x.dtype
    assert gemm_fused.numel() == 2 * x.numel()
    px, pa = gemm_fused.data_ptr(), gemm_fused.data_ptr() + block_width * x.element_size()
    ldx, gbw = 2 * width, block_width
    keep = (gemm_fused,)
  else:
    assert gemm_x.shape == x.shape and gemm_a.shape == x.shape
    assert gemm_x.dtype == x.dtype and gemm_a.dtype == x.dtype

    def rows(t):
      if t.stride(2) == 1 and t.stride(0) == steps * t.stride(1):
        return t, t.stride(1)
      t = t.contiguous()
      return t, width

    gemm_x, ldx = rows(gemm_x)
    gemm_a, lda = rows(gemm_a)
    if ldx != lda:
      gemm_x, gemm_a, ldx = gemm_x.contiguous(), gemm_a.contiguous(), width
    px, pa, gbw = gemm_x.data_ptr(), gemm_a.data_ptr(), 0
    keep = (gemm_x, gemm_a)
  seg, is64, stride = _seg_args(segment_pos, bsz, steps)
  y = torch.empty_like(x) if out is None else out
  last_h = None
  if return_cache:
    last_h = (torch.empty((bsz, width), dtype=torch.float32, device=x.device)
              if last_h_out is None else last_h_out)
  ws = scan_workspace(x.device, bsz, steps, width, x.dtype)
  bx = None if bias_x is None else bias_x.contiguous().view(-1)
  ba = None if bias_a is None else bias_a.contiguous().view(-1)
  h0c = None if h0 is None else h0.contiguous()
  with _on_device(x.device):
    rc = load().cg_rglru_fwd(x.data_ptr(), px, pa, ldx, gbw, _ptr(bx), _ptr(ba),
                             a_param.contiguous().data_ptr(), seg.data_ptr(), is64,
                             stride, _ptr(h0c), y.data_ptr(), _ptr(last_h),
                             ws.data_ptr(), ws.numel(), bsz, steps, width,
                             dtype_code(x.dtype), arith_mode, _stream(x))
  del keep
  _check(rc, "cg_rglru_fwd")
  launch_count += 2   # prologue (ticket/epoch/softplus) + scan kernel
  return y, last_h


def rnn_scan_fwd(x, a, reset, h0=None, arith_mode=ARITH_REFERENCE):
  global launch_count
  _require_cuda(x, a, reset, h0)
  bsz, steps, width = x.shape
  x, a = x.contiguous(), a.contiguous()
  rs = reset.to(torch.uint8).contiguous()
  assert rs.shape == (bsz, steps)
  y = torch.empty_like(x)
  h_last = torch.empty((bsz, width), dtype=torch.float32, device=x.device)
  ws = scan_workspace(x.device, bsz, steps, width, x.dtype)
  h0c = None if h0 is None else h0.contiguous()
  with _on_device(x.device):
    rc = load().cg_rnn_scan_fwd(x.data_ptr(), a.data_ptr(), rs.data_ptr(),
                                _ptr(h0c), y.data_ptr(), h_last.data_ptr(),
                                ws.data_ptr(), ws.numel(), bsz, steps, width,
                                dtype_code(x.dtype), arith_mode, _stream(x))
  _check(rc, "cg_rnn_scan_fwd")
  launch_count += 1 if (arith_mode & ARITH_STRICT) else 2
  return y, h_last


def rglru_gates_fwd(x, pre_x, pre_a, a_param, reset):
  """Gate math of the RG-LRU for the training path (cg_rglru_gates_fwd):
  returns ``(a, normalized_x)``, the inputs of ``rnn_scan``."""
  global launch_count
  _require_cuda(x, pre_x, pre_a, a_param, reset)
  bsz, steps, width = x.shape
  assert pre_x.shape == x.shape and pre_a.shape == x.shape
  assert pre_x.dtype == x.dtype and pre_a.dtype == x.dtype and a_param.dtype == x.dtype
  x, pre_x, pre_a, ap = x.contiguous(), pre_x.contiguous(), pre_a.contiguous(), a_param.contiguous()
  rs = reset.to(torch.uint8).contiguous()
  a, nx = torch.empty_like(x), torch.empty_like(x)
  with _on_device(x.device):
    rc = load().cg_rglru_gates_fwd(x.data_ptr(), pre_x.data_ptr(), pre_a.data_ptr(), ap.data_ptr(),
                                   rs.data_ptr(), a.data_ptr(), nx.data_ptr(), bsz, steps, width,
                                   dtype_code(x.dtype), _stream(x))
  _check(rc, "cg_rglru_gates_fwd")
  launch_count += 1
  return a, nx


def rglru_gates_bwd(x, pre_x, pre_a, a_param, reset, d_nx, d_a):
  """Backward of ``rglru_gates_fwd`` (cg_rglru_gates_bwd): returns
  ``(dx, d_pre_x, d_pre_a, d_a_param)``."""
  global launch_count
  _require_cuda(x, pre_x, pre_a, a_param, reset, d_nx, d_a)
  bsz, steps, width = x.shape
  x, pre_x, pre_a, ap = x.contiguous(), pre_x.contiguous(), pre_a.contiguous(), a_param.contiguous()
  d_nx, d_a = d_nx.contiguous(), d_a.contiguous()
  assert d_nx.dtype == x.dtype and d_a.dtype == x.dtype
  rs = reset.to(torch.uint8).contiguous()
  dx, dpx, dpa = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
  dap = torch.empty_like(ap)
  ws = torch.empty(load().cg_rglru_gates_bwd_workspace_bytes(bsz, steps, width), dtype=torch.uint8,
                   device=x.device)
  with _on_device(x.device):
    rc = load().cg_rglru_gates_bwd(x.data_ptr(), pre_x.data_ptr(), pre_a.data_ptr(), ap.data_ptr(),
                                   rs.data_ptr(), d_nx.data_ptr(), d_a.data_ptr(), dx.data_ptr(),
                                   dpx.data_ptr(), dpa.data_ptr(), dap.data_ptr(), ws.data_ptr(),
                                   ws.numel(), bsz, steps, width, dtype_code(x.dtype), _stream(x))
  _check(rc, "cg_rglru_gates_bwd")
  launch_count += 2
  return dx, dpx, dpa, dap


def rnn_scan_bwd(gy, g_last, a, h, reset, h0=None, need_dh0=True):
  """Backward of ``rnn_scan`` (cg_rnn_scan_bwd): returns ``(dx, da, dh0 | None)``.

  ``gy`` grad of y, ``g_last`` fp32 grad of last_h or None, ``a`` / ``reset`` /
  ``h0`` the forward's inputs, ``h`` the forward's output y."""
  global launch_count
  _require_cuda(gy, a, h, reset, h0, g_last)
  bsz, steps, width = gy.shape
  assert a.shape == gy.shape and h.shape == gy.shape and a.dtype == gy.dtype and h.dtype == gy.dtype
  gy, a, h = gy.contiguous(), a.contiguous(), h.contiguous()
  rs = reset.to(torch.uint8).contiguous()
  assert rs.shape == (bsz, steps)
  glc = None if g_last is None else g_last.to(torch.float32).contiguous()
  h0c = None if h0 is None else h0.contiguous()
  dx, da = torch.empty_like(gy), torch.empty_like(gy)
  dh0 = torch.empty((bsz, width), dtype=torch.float32, device=gy.device) if need_dh0 else None
  ws = scan_workspace(gy.device, bsz, steps, width, gy.dtype)
  with _on_device(gy.device):
    rc = load().cg_rnn_scan_bwd(gy.data_ptr(), _ptr(glc), a.data_ptr(), h.data_ptr(), rs.data_ptr(),
                                _ptr(h0c), dx.data_ptr(), da.data_ptr(), _ptr(dh0), ws.data_ptr(),
                                ws.numel(), bsz, steps, width, dtype_code(gy.dtype), _stream(gy))
  _check(rc, "cg_rnn_scan_bwd")
  launch_count += 2
  return dx, da, dh0


# ---------------------------------------------------------------------------
# Fused tensor-core path: gate GEMMs + gate math + scan in one kernel
# ---------------------------------------------------------------------------
def fused_supported(width: int, heads: int, dtype) -> bool:
  """True if cg_rglru_fused_fwd takes this (E, H, dtype)."""
  if dtype != torch.bfloat16:
    return False
  return bool(load().cg_rglru_fused_supported(width, heads, DTYPE_BF16))


def pack_gate_weights(wx: torch.Tensor, wa: torch.Tensor) -> torch.Tensor:
  """[H, bw, bw] x 2 -> the packed shared-memory image the fused kernel reads."""
  global launch_count
  _require_cuda(wx, wa)
  heads, bw, bw2 = wx.shape
  assert bw == bw2 and wa.shape == wx.shape and wa.dtype == wx.dtype
  width = heads * bw
  nbytes = load().cg_rglru_gate_pack_bytes(width, heads)
  assert nbytes > 0, "shape not supported by the fused path"
  out = torch.empty(nbytes, dtype=torch.uint8, device=wx.device)
  wx, wa = wx.detach().contiguous(), wa.detach().contiguous()
  with _on_device(wx.device):
    rc = load().cg_rglru_pack_gate_weights(wx.data_ptr(), wa.data_ptr(), out.data_ptr(),
                                           width, heads, dtype_code(wx.dtype), _stream(wx))
  _check(rc, "cg_rglru_pack_gate_weights")
  launch_count += 1
  return out


def fused_workspace(device, batch: int, steps: int, width: int) -> torch.Tensor:
  """Scratch of the fused tensor-core kernels for the current stream (grow-only)."""
  nbytes = load().cg_rglru_fused_workspace_bytes(batch, steps, width)
  return _grow_only_workspace("fused", device, nbytes)


def rglru_fused_fwd(x, wpack, bias_x, bias_a, a_param, segment_pos, heads, h0=None,
                    return_cache=True, arith_mode=ARITH_FAST, out=None, debug=False,
                    workspace=None, last_h_out=None, gate_mul=None):
  """RGLRU.forward (gate GEMMs included) on the fused tcgen05 kernel.

  ``gate_mul`` ([B,T,E], optional): return ``round(y * gate_mul)`` -- the gating
  product of ``RecurrentBlock.forward`` (reference modules.py:651) folded in.
  ``debug`` (tests): also return the pre-activation / x taps (experimental ABI)."""
  res = recurrent_prefill_fwd(x, None, None, wpack, bias_x, bias_a, a_param, segment_pos, heads, h0=h0,
                              return_cache=return_cache, arith_mode=arith_mode, out=out, debug=debug,
                              workspace=workspace, last_h_out=last_h_out, gate_mul=gate_mul)
  return (res[0], res[2]) + tuple(res[3:])


def recurrent_prefill_fwd(x, conv_w, conv_b, wpack, bias_x, bias_a, a_param, segment_pos, heads,
                          h0=None, return_cache=True, mask_mode=MASK_FORK, arith_mode=ARITH_FAST,
                          out=None, conv_cache_out=None, last_h_out=None, gate_mul=None, debug=False,
                          workspace=None):
  """``Conv1D.forward -> RGLRU.forward`` (prefill) in ONE fused kernel
  (``cg_recurrent_prefill_fwd``): ``x`` is the convolution's INPUT, the temporal
  convolution runs inside the tcgen05 kernel.  With ``conv_w is None`` ``x`` is
  the convolution's output and the call is ``cg_rglru_fused_fwd``.

  Returns ``(y, conv_cache | None, last_h | None[, debug taps])``."""
  global launch_count
  _require_cuda(x, conv_w, conv_b, wpack, bias_x, bias_a, a_param, segment_pos, h0, gate_mul)
  bsz, steps, width = x.shape
  assert x.dtype == torch.bfloat16 and a_param.dtype == x.dtype
  conv = conv_w is not None
  if conv:
    assert conv_w.shape == (4, width) and conv_b.shape == (width,), "fused prefill: temporal width 4"
    assert conv_w.dtype == x.dtype and conv_b.dtype == x.dtype
    conv_w, conv_b = conv_w.contiguous(), conv_b.contiguous()
  if gate_mul is not None:
    assert gate_mul.shape == x.shape and gate_mul.dtype == x.dtype
    gate_mul = gate_mul.contiguous()
  assert h0 is None or h0.dtype == torch.float32, "layers.py:170"
  x = x.contiguous()
  seg, is64, stride = _seg_args(segment_pos, bsz, steps)
  y = torch.empty_like(x) if out is None else out
  assert y.shape == x.shape and y.dtype == x.dtype and y.is_contiguous()
  last_h = conv_cache = None
  if return_cache:
    last_h = (torch.empty((bsz, width), dtype=torch.float32, device=x.device)
              if last_h_out is None else last_h_out)
    if conv:
      conv_cache = (torch.empty((bsz, 3, width), dtype=x.dtype, device=x.device)
                    if conv_cache_out is None else conv_cache_out)
      assert conv_cache.shape == (bsz, 3, width) and conv_cache.dtype == x.dtype and conv_cache.is_contiguous()
  ws = fused_workspace(x.device, bsz, steps, width) if workspace is None else workspace
  bx = None if bias_x is None else bias_x.contiguous().view(-1)
  ba = None if bias_a is None else bias_a.contiguous().view(-1)
  h0c = None if h0 is None else h0.contiguous()
  ap = a_param.contiguous()
  lib = load()
  with _on_device(x.device):
    if debug:
      assert gate_mul is None
      dbg = torch.zeros((3, bsz, steps, width), dtype=x.dtype, device=x.device)
      rc = lib.cg_recurrent_prefill_debug(x.data_ptr(), _ptr(conv_w), _ptr(conv_b), wpack.data_ptr(),
                                          _ptr(bx), _ptr(ba), ap.data_ptr(), seg.data_ptr(), is64, stride,
                                          _ptr(h0c), y.data_ptr(), _ptr(conv_cache), _ptr(last_h),
                                          ws.data_ptr(), ws.numel(), bsz, steps, width, heads,
                                          dtype_code(x.dtype), mask_mode, arith_mode, dbg.data_ptr(), _stream(x))
      what = "cg_recurrent_prefill_debug"
    elif conv:
      rc = lib.cg_recurrent_prefill_fwd(x.data_ptr(), conv_w.data_ptr(), conv_b.data_ptr(), wpack.data_ptr(),
                                        _ptr(bx), _ptr(ba), ap.data_ptr(), seg.data_ptr(), is64, stride,
                                        _ptr(h0c), _ptr(gate_mul), y.data_ptr(), _ptr(conv_cache), _ptr(last_h),
                                        ws.data_ptr(), ws.numel(), bsz, steps, width, heads, 4,
                                        dtype_code(x.dtype), mask_mode, arith_mode, _stream(x))
      what = "cg_recurrent_prefill_fwd"
    else:
      rc = lib.cg_rglru_fused_fwd(x.data_ptr(), wpack.data_ptr(), _ptr(bx), _ptr(ba), ap.data_ptr(),
                                  seg.data_ptr(), is64, stride, _ptr(h0c), y.data_ptr(), _ptr(last_h),
                                  ws.data_ptr(), ws.numel(), bsz, steps, width, heads,
                                  dtype_code(x.dtype), arith_mode, _ptr(gate_mul), _stream(x))
      what = "cg_rglru_fused_fwd"
  _check(rc, what)
  launch_count += 2   # prologue + fused kernel
  if debug:
    return y, conv_cache, last_h, dbg
  return y, conv_cache, last_h


# ---------------------------------------------------------------------------
# Fused decode step (conv step + gate GEMVs + gates + h = a*h0 + x~, one launch)
# ---------------------------------------------------------------------------
def decode_supported(width: int, heads: int, temporal_width: int, dtype) -> bool:
  if dtype != torch.bfloat16:
    return False
  return bool(load().cg_recurrent_decode_supported(width, heads, temporal_width, DTYPE_BF16))


def recurrent_decode_step(x, conv_w, conv_b, conv_cache, wx, wa, bias_x, bias_a, a_param,
                          segment_pos, h0=None, gate_mul=None, return_cache=True,
                          arith_mode=ARITH_FAST):
  """One decode step of Conv1D -> RG-LRU.  Returns ``(y, new_conv_cache | None, last_h | None)``."""
  global launch_count
  _require_cuda(x, conv_w, conv_b, conv_cache, wx, wa, a_param, segment_pos, h0, gate_mul)
  bsz, steps, width = x.shape
  heads, bw, _ = wx.shape
  assert steps == 1, "layers.py:566: decode takes exactly one token"
  assert conv_cache.shape == (bsz, conv_w.shape[0] - 1, width), "layers.py:565"
  assert h0 is None or h0.dtype == torch.float32, "layers.py:170"
  x, conv_cache = x.contiguous(), conv_cache.contiguous()
  wx, wa = wx.contiguous(), wa.contiguous()
  seg, is64, stride = _seg_args(segment_pos, bsz, 1)
  y = torch.empty_like(x)
  new_cache = torch.empty_like(conv_cache) if return_cache else None
  last_h = (torch.empty((bsz, width), dtype=torch.float32, device=x.device)
            if return_cache else None)
  bx = None if bias_x is None else bias_x.contiguous().view(-1)
  ba = None if bias_a is None else bias_a.contiguous().view(-1)
  h0c = None if h0 is None else h0.contiguous()
  gm = None if gate_mul is None else gate_mul.contiguous()
  with _on_device(x.device):
    rc = load().cg_recurrent_decode_step(
        x.data_ptr(), conv_w.contiguous().data_ptr(), conv_b.contiguous().data_ptr(),
        conv_cache.data_ptr(), dtype_code(conv_cache.dtype), wx.data_ptr(), wa.data_ptr(),
        _ptr(bx), _ptr(ba), a_param.contiguous().data_ptr(), seg.data_ptr(), is64, stride,
        _ptr(h0c), _ptr(gm), y.data_ptr(), _ptr(new_cache), _ptr(last_h), bsz, width, heads,
        conv_w.shape[0], dtype_code(x.dtype), arith_mode & ARITH_FAST, _stream(x))
  _check(rc, "cg_recurrent_decode_step")
  launch_count += 1
  return y, new_cache, last_h
