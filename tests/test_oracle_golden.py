"""Pins both CPU oracles to golden vectors produced by the unmodified reference.

`oracle/torch_port.py` issues the same ATen ops as the reference and must be
bit-exact on the same torch build; `oracle/rglru_oracle.c` is an independent
scalar restatement (libm instead of SLEEF) and must be bit-exact wherever no
transcendental is involved and within a few fp32 ulp elsewhere.
"""
import pytest
import torch

from oracle import c_oracle, torch_port
from tests.golden import fixture_io
from tests.helpers import (normwise, assert_bitexact, assert_close_bf16, assert_close_f32,
                           identical_fraction)


def _close(a, b, what, min_identical=0.995):
  if a.dtype == torch.bfloat16:
    assert_close_bf16(a, b, what, min_identical=min_identical)
  else:
    assert_close_f32(a, b, what)


# ---------------------------------------------------------------- rnn_scan
@pytest.mark.parametrize("case", fixture_io.cases("rnn_scan_"))
def test_rnn_scan(case):
  g = fixture_io.load(case)
  h0 = g.get("h0")
  y, h = torch_port.rnn_scan(g["x"], g["a"], g["reset"], h0)
  assert_bitexact(y, g["y"], case + " port y")
  assert_bitexact(h, g["h_last"], case + " port h")
  y, h = c_oracle.rnn_scan(g["x"], g["a"], g["reset"], h0)
  assert_bitexact(y, g["y"], case + " C y")
  assert_bitexact(h, g["h_last"], case + " C h")


@pytest.mark.parametrize("case", fixture_io.cases("grad_rnn_scan_"))
def test_rnn_scan_gradients(case):
  """Autograd through the port's loop == autograd through the reference's
  (the oracle of the backward scan kernel)."""
  g = fixture_io.load(case)
  x = g["x"].clone().requires_grad_()
  a = g["a"].clone().requires_grad_()
  h0 = g["h0"].clone().requires_grad_() if "h0" in g else None
  with torch.enable_grad():
    y, h = torch_port.rnn_scan(x, a, g["reset"], h0)
    torch.autograd.backward([y, h], [g["gy"], g["gh"]])
  assert_bitexact(x.grad, g["dx"], case + " dx")
  if "da" in g:
    assert_bitexact(a.grad, g["da"], case + " da")
  else:                                   # T == 1 without h0 returns x itself (:177-178): a is unused
    assert a.grad is None
  if h0 is not None:
    assert_bitexact(h0.grad, g["dh0"], case + " dh0")


# ------------------------------------------------------------------ conv1d
@pytest.mark.parametrize("case", fixture_io.cases("conv1d_"))
def test_conv1d(case):
  g = fixture_io.load(case)
  for name, fn in (("port", torch_port.conv1d_forward),
                   ("C", c_oracle.conv1d_forward)):
    if "x" in g:
      x_before = g["x"].clone()
      y, cache = fn(g["w"], g["b"], g["x"], g["seg"])
      assert torch.equal(g["x"], x_before), "oracle must not mutate its input"
      assert_bitexact(y, g["y"], f"{case} {name} y")
      assert_bitexact(cache, g["cache"], f"{case} {name} cache")
    else:
      cache = g["cache_in"]
    for i in range(2):
      if f"step{i}_x" not in g:
        break
      ys, cache = fn(g["w"], g["b"], g[f"step{i}_x"], None, cache)
      assert_bitexact(ys, g[f"step{i}_y"], f"{case} {name} step{i} y")
      assert_bitexact(cache, g[f"step{i}_cache"], f"{case} {name} step{i} cache")


# ------------------------------------------------------------------- rglru
@pytest.mark.parametrize("case", fixture_io.cases("rglru_"))
def test_rglru(case):
  g = fixture_io.load(case)
  p = torch_port.RGLRUParams(g["a_param"], g["input_gate_w"], g["input_gate_b"],
                             g["a_gate_w"], g["a_gate_b"])
  # torch port: same ATen ops -> identical on the generating torch build; allow
  # rare SLEEF/ISA differences on another host (>= 99.9 % identical).
  y, h = torch_port.rglru_forward(p, g["x"], g["seg"])
  assert identical_fraction(y, g["y"]) >= 0.999, case
  _close(y, g["y"], case + " port y")
  _close(h, g["last_h"], case + " port last_h")
  cache = h
  for i in range(2):
    ys, cache = torch_port.rglru_forward(p, g[f"step{i}_x"],
                                         g["seg"][:, -1:] + 1 + i, cache)
    _close(ys, g[f"step{i}_y"], f"{case} port step{i}")
    _close(cache, g[f"step{i}_last_h"], f"{case} port step{i} h")
  # kernel-boundary form (same pre-activations as the reference computed)
  y, h = torch_port.rglru_from_preacts(g["x"], g["pre_x"], g["pre_a"],
                                       g["a_param"], g["seg"])
  assert identical_fraction(y, g["y"]) >= 0.999, case
  # C oracle from the reference's own pre-activations
  y, h = c_oracle.rglru_from_preacts(g["x"], g["pre_x"], g["pre_a"],
                                     g["a_param"], g["seg"])
  _close(y, g["y"], case + " C y")
  _close(h, g["last_h"], case + " C last_h")
  cache = h
  for i in range(2):
    ys, cache = c_oracle.rglru_from_preacts(
        g[f"step{i}_x"], g[f"step{i}_pre_x"], g[f"step{i}_pre_a"],
        g["a_param"], g["seg"][:, -1:] + 1 + i, cache)
    _close(ys, g[f"step{i}_y"], f"{case} C step{i}")
    _close(cache, g[f"step{i}_last_h"], f"{case} C step{i} h")
  # C block-diagonal GEMM + bias split (what the CUDA ABI consumes)
  zeros = torch.zeros_like(g["input_gate_b"])
  gx = c_oracle.block_diagonal_linear(g["x"], g["input_gate_w"], zeros)
  ga = c_oracle.block_diagonal_linear(g["x"], g["a_gate_w"], zeros)
  y, h = c_oracle.rglru_from_preacts(
      g["x"], gx, ga, g["a_param"], g["seg"], bias_x=g["input_gate_b"],
      bias_a=g["a_gate_b"])
  _close(y, g["y"], case + " C full y", min_identical=0.97)
  if g["x"].dtype == torch.bfloat16:
    # a different GEMM summation order flips a few bf16 pre-activations; the
    # fp32 state then differs at bf16 resolution, not at fp32 resolution
    assert normwise(h, g["last_h"]) <= 1e-2, case + " C full last_h"
  else:
    _close(h, g["last_h"], case + " C full last_h")


# --------------------------------------------------------- recurrent block
@pytest.mark.parametrize("case", fixture_io.cases("recurrent_block_"))
def test_recurrent_block(case):
  g = fixture_io.load(case)
  P = lambda k: g["param." + k]
  p = torch_port.RecurrentBlockParams(
      P("linear_y.weight"), P("linear_y.bias"), P("linear_x.weight"),
      P("linear_x.bias"), P("linear_out.weight"), P("linear_out.bias"),
      P("conv_1d.w"), P("conv_1d.b"),
      torch_port.RGLRUParams(P("rg_lru.a_param"), P("rg_lru.input_gate.w"),
                             P("rg_lru.input_gate.b"), P("rg_lru.a_gate.w"),
                             P("rg_lru.a_gate.b")))
  y, cache = torch_port.recurrent_block_forward(p, g["x"], g["seg"])
  _close(y, g["y"], case + " y")
  _close(cache[0], g["rg_lru_state"], case + " lru state")
  assert_bitexact(cache[1], g["conv1d_state"], case + " conv state")
  for i in range(2):
    ys, cache = torch_port.recurrent_block_forward(
        p, g[f"step{i}_x"], g["seg"][:, -1:] + 1 + i, cache)
    _close(ys, g[f"step{i}_y"], f"{case} step{i} y")
    _close(cache[0], g[f"step{i}_rg_lru_state"], f"{case} step{i} lru")
    assert_bitexact(cache[1], g[f"step{i}_conv1d_state"], f"{case} step{i} conv")


# ------------------------------------------------- griffin-tiny hot-path taps
def test_griffin_tiny_hot_path():
  g = fixture_io.load("griffin_tiny_f32_t128")
  for blk in (0, 1):
    P = lambda k: g[f"blk{blk}_param.{k}"]
    xc, conv_state = c_oracle.conv1d_forward(P("conv_1d.w"), P("conv_1d.b"),
                                             g[f"blk{blk}_conv_in"], g["seg"])
    assert_bitexact(conv_state, g[f"blk{blk}_conv1d_state"], "conv state")
    y, h = c_oracle.rglru_forward(
        P("rg_lru.a_param"), P("rg_lru.input_gate.w"), P("rg_lru.input_gate.b"),
        P("rg_lru.a_gate.w"), P("rg_lru.a_gate.b"), xc, g["seg"])
    assert_close_f32(y, g[f"blk{blk}_rglru_out"], f"blk{blk} rglru out")
    assert_close_f32(h, g[f"blk{blk}_rg_lru_state"], f"blk{blk} state")
