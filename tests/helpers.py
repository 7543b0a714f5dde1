"""Shared comparison helpers for the parity tests."""
from __future__ import annotations

import torch


def normwise(a: torch.Tensor, b: torch.Tensor) -> float:
  """max|a-b| / max|b| in fp64 (the parity metric of SURVEY.md section 8c)."""
  a64, b64 = a.double(), b.double()
  denom = b64.abs().max().item()
  return (a64 - b64).abs().max().item() / (denom if denom > 0 else 1.0)


def identical_fraction(a: torch.Tensor, b: torch.Tensor) -> float:
  """Fraction of elements that are bit-identical (+0 == -0 counts as equal)."""
  return (a == b).double().mean().item() if a.numel() else 1.0


def assert_bitexact(a, b, what=""):
  assert a.dtype == b.dtype and a.shape == b.shape, (what, a.dtype, b.dtype,
                                                     a.shape, b.shape)
  frac = identical_fraction(a, b)
  assert frac == 1.0, f"{what}: only {frac:.6f} identical, normwise {normwise(a, b):.3e}"


def assert_close_bf16(a, b, what="", min_identical=0.99):
  """bf16 parity: the reference's own tolerance (layers_test.py:131) + flips."""
  assert a.dtype == b.dtype == torch.bfloat16, (what, a.dtype, b.dtype)
  torch.testing.assert_close(a.float(), b.float(), rtol=1e-2, atol=3e-2,
                             msg=lambda m: f"{what}: {m}")
  frac = identical_fraction(a, b)
  assert frac >= min_identical, f"{what}: only {frac:.5f} bit-identical"


def assert_close_f32(a, b, what="", tol=1e-5):
  """fp32 parity: <= 1e-5 normwise (north star) and elementwise with that atol."""
  assert a.dtype == b.dtype == torch.float32, (what, a.dtype, b.dtype)
  nw = normwise(a, b)
  assert nw <= tol, f"{what}: normwise {nw:.3e} > {tol}"
  scale = b.abs().max().item()
  torch.testing.assert_close(a, b, rtol=tol, atol=tol * max(scale, 1e-30),
                             msg=lambda m: f"{what}: {m}")
