"""ORACLE (test infrastructure) -- ctypes front-end of ``oracle/rglru_oracle.c``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg
may import this.  Tensors are CPU ``torch`` tensors (fp32 or bf16, contiguous);
bf16 is handed to C as raw ``uint16`` words.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liborc.so")
_lib = None

F32, BF16 = 0, 1


def build(force: bool = False) -> str:
  """Compiles the oracle with gcc (``oracle/Makefile``). Returns the .so path."""
  src = os.path.join(_HERE, "rglru_oracle.c")
  stale = (not os.path.exists(_LIB_PATH) or
           os.path.getmtime(_LIB_PATH) < os.path.getmtime(src))
  if force or stale:
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
  return _LIB_PATH


def lib():
  global _lib
  if _lib is None:
    if not os.path.exists(_LIB_PATH):
      build()
    _lib = ctypes.CDLL(_LIB_PATH)
    for name in ("orc_rnn_scan", "orc_block_diag_linear",
                 "orc_rglru_from_preacts", "orc_conv1d_fwd",
                 "orc_abi_version"):
      getattr(_lib, name).restype = ctypes.c_int
  return _lib


def _dt(t: torch.Tensor) -> int:
  if t.dtype == torch.float32:
    return F32
  if t.dtype == torch.bfloat16:
    return BF16
  raise TypeError(f"oracle supports fp32/bf16, got {t.dtype}")


def _p(t):
  if t is None:
    return ctypes.c_void_p(0)
  assert t.device.type == "cpu" and t.is_contiguous()
  return ctypes.c_void_p(t.data_ptr())


def _seg(segment_pos, batch):
  seg = segment_pos.to(torch.int64)
  if seg.ndim == 1:
    seg = seg[None, :]
  seg = seg.contiguous()
  stride = 0 if (seg.shape[0] == 1 and batch > 1) else seg.shape[1]
  return seg, stride


def rnn_scan(x, a, reset, h0):
  """layers.py:146-199 -> (y, h_last fp32)."""
  x, a = x.contiguous(), a.contiguous()
  bsz, steps, width = x.shape
  rs = reset.to(torch.uint8).contiguous()
  y = torch.empty_like(x)
  h_last = torch.empty((bsz, width), dtype=torch.float32)
  rc = lib().orc_rnn_scan(_p(x), _p(a), _p(rs), _p(h0), _p(y), _p(h_last),
                          bsz, steps, width, _dt(x))
  assert rc == 0, rc
  return y, h_last


def block_diagonal_linear(x, w, b):
  """layers.py:133-142."""
  x, w, b = x.contiguous(), w.contiguous(), b.contiguous()
  heads, bw, _ = w.shape
  y = torch.empty_like(x)
  n = x.numel() // (heads * bw)
  rc = lib().orc_block_diag_linear(_p(x), _p(w), _p(b), _p(y),
                                   ctypes.c_int64(n), heads, bw, _dt(x))
  assert rc == 0, rc
  return y


def rglru_from_preacts(x, gemm_x, gemm_a, a_param, segment_pos, h0=None,
                       bias_x=None, bias_a=None, arith_mode=0):
  """layers.py:345-375 from the gate GEMM outputs -> (y, last_h fp32)."""
  x = x.contiguous()
  bsz, steps, width = x.shape
  seg, stride = _seg(segment_pos, bsz)
  y = torch.empty_like(x)
  last_h = torch.empty((bsz, width), dtype=torch.float32)
  rc = lib().orc_rglru_from_preacts(
      _p(x), _p(gemm_x.contiguous()), _p(gemm_a.contiguous()),
      _p(None if bias_x is None else bias_x.contiguous().view(-1)),
      _p(None if bias_a is None else bias_a.contiguous().view(-1)),
      _p(a_param.contiguous()), _p(seg), ctypes.c_int64(stride), _p(h0),
      _p(y), _p(last_h), bsz, steps, width, _dt(x), arith_mode)
  assert rc == 0, rc
  return y, last_h


def rglru_forward(a_param, input_gate_w, input_gate_b, a_gate_w, a_gate_b, x,
                  segment_pos, cache=None, arith_mode=0):
  """layers.py:322-375 (full layer incl. the two block-diagonal gates)."""
  pre_x = block_diagonal_linear(x, input_gate_w, input_gate_b)
  pre_a = block_diagonal_linear(x, a_gate_w, a_gate_b)
  return rglru_from_preacts(x, pre_x, pre_a, a_param, segment_pos, cache,
                            arith_mode=arith_mode)


def conv1d_forward(w, b, x, segment_pos, cache=None, mask_mode=0,
                   arith_mode=0):
  """layers.py:458-546 -> (y, new_cache)."""
  x, w, b = x.contiguous(), w.contiguous(), b.contiguous()
  bsz, steps, width = x.shape
  tw = w.shape[0]
  y = torch.empty_like(x)
  if cache is not None:
    assert steps == 1, "layers.py:566"
    cache = cache.contiguous()
    new_cache = torch.empty_like(cache)
    rc = lib().orc_conv1d_fwd(_p(x), _p(w), _p(b), _p(None),
                              ctypes.c_int64(0), _p(cache), _dt(cache), _p(y),
                              _p(new_cache), bsz, steps, width, tw, _dt(x),
                              mask_mode, arith_mode)
  else:
    seg, stride = _seg(segment_pos, bsz)
    new_cache = torch.empty((bsz, tw - 1, width), dtype=x.dtype)
    rc = lib().orc_conv1d_fwd(_p(x), _p(w), _p(b), _p(seg),
                              ctypes.c_int64(stride), _p(None), _dt(x), _p(y),
                              _p(new_cache), bsz, steps, width, tw, _dt(x),
                              mask_mode, arith_mode)
  assert rc == 0, rc
  return y, new_cache
