"""GPU parity tests: the sm_100a kernels (through the C ABI) against

  * the golden vectors produced by the unmodified reference (tests/golden/),
  * the CPU oracles on fresh seeded inputs (ragged sizes, resets, h0, padding),
  * the strict sequential device kernel at BASELINE.json's full sizes, plus
    size-independent properties (chunked continuation, reset isolation,
    run-to-run determinism).

Tolerances (north star): fp32 <= 1e-5 normwise; bf16 within the reference's
own test tolerance rtol 1e-2 / atol 3e-2 (recurrentgemma/torch/layers_test.py
:131) *and* a minimum fraction of bit-identical elements; Conv1D and the
strict scan are bit-exact.
"""
import pytest
import torch

from tests.golden import fixture_io
from tests.helpers import (assert_bitexact, assert_close_bf16, assert_close_f32,
                           identical_fraction, normwise)

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
REF, FP32, FAST, STRICT = 0, 1, 2, 4


def _abi():
  from cadence_gemma_b200 import _abi as abi
  abi.load()
  return abi


def cu(t):
  return None if t is None else t.to(DEV)


def _close(a, b, what, min_identical=0.98):
  a, b = a.cpu(), b.cpu()
  if a.dtype == torch.bfloat16:
    assert_close_bf16(a, b, what, min_identical=min_identical)
  else:
    assert_close_f32(a, b, what)


# ------------------------------------------------------------------ rnn_scan
@pytest.mark.parametrize("case", fixture_io.cases("rnn_scan_"))
def test_rnn_scan_golden(case):
  abi = _abi()
  g = fixture_io.load(case)
  x, a, rs, h0 = cu(g["x"]), cu(g["a"]), cu(g["reset"]), cu(g.get("h0"))
  y, h = abi.rnn_scan_fwd(x, a, rs, h0, arith_mode=STRICT)
  assert_bitexact(y.cpu(), g["y"], case + " strict y")
  assert_bitexact(h.cpu(), g["h_last"], case + " strict h")
  y, h = abi.rnn_scan_fwd(x, a, rs, h0, arith_mode=REF)
  _close(y, g["y"], case + " y", min_identical=0.995)
  assert_close_f32(h.cpu(), g["h_last"], case + " h")


def test_rnn_scan_api_matches_reference_contract():
  import cadence_gemma_b200 as cg
  x = torch.randn(2, 5, 64, device=DEV, dtype=torch.bfloat16)
  a = torch.rand_like(x)
  rs = torch.zeros(2, 5, dtype=torch.bool, device=DEV)
  with pytest.raises(AssertionError):      # layers.py:170: h0 must be fp32
    cg.rnn_scan(x, a, rs, torch.zeros(2, 64, device=DEV, dtype=torch.bfloat16))
  y, h = cg.rnn_scan(x, a, rs, None)
  assert y.dtype == torch.bfloat16 and h.dtype == torch.float32
  assert y.shape == x.shape and h.shape == (2, 64)
  with pytest.raises(RuntimeError):        # no CPU fallback
    cg.rnn_scan(x.cpu(), a.cpu(), rs.cpu(), None)


# ------------------------------------------------- gradients
def _grad_close(got, want, what, bf16_identical=0.97):
  got, want = got.cpu(), want.cpu()
  assert got.dtype == want.dtype and got.shape == want.shape, (what, got.dtype, want.dtype)
  if got.dtype == torch.bfloat16:
    assert_close_bf16(got, want, what, min_identical=bf16_identical)
  else:
    assert_close_f32(got, want, what)


@pytest.mark.parametrize("case", fixture_io.cases("grad_rnn_scan_"))
def test_rnn_scan_backward_golden(case):
  """cg_rnn_scan_bwd against the gradients autograd derives from the reference
  loop.  fp32: 1e-5 normwise.  bf16: dx to bf16 rounding; da is formed from the
  bf16 forward output instead of the reference's fp32 h_t (documented in
  include/cadence_b200.h), hence the lower bit-identical floor."""
  import cadence_gemma_b200 as cg
  g = fixture_io.load(case)
  x = cu(g["x"]).requires_grad_()
  a = cu(g["a"]).requires_grad_()
  h0 = cu(g["h0"]).requires_grad_() if "h0" in g else None
  with torch.enable_grad():
    y, h = cg.rnn_scan(x, a, cu(g["reset"]), h0)
    torch.autograd.backward([y, h], [cu(g["gy"]), cu(g["gh"])])
  _close(y.detach(), g["y"], case + " y", min_identical=0.995)
  _grad_close(x.grad, g["dx"], case + " dx", 0.995)
  if "da" in g:
    _grad_close(a.grad, g["da"], case + " da", 0.5)
  else:                                   # T == 1 without h0 (:177-178): a does not reach the output
    assert a.grad is None or not a.grad.any()
  if h0 is not None:
    assert_close_f32(h0.grad.cpu(), g["dh0"], case + " dh0")
  # direct ABI call, grad of last_h absent and dh0 not wanted
  abi = _abi()
  dx, da, dh0 = abi.rnn_scan_bwd(cu(g["gy"]), None, a.detach(), y.detach(), cu(g["reset"]),
                                 None if h0 is None else h0.detach(), need_dh0=False)
  assert dh0 is None and torch.isfinite(dx.float()).all() and torch.isfinite(da.float()).all()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_rnn_scan_backward_full_size_vs_sequential(dtype):
  """Config-2 sized backward scan (many chunks, look-back across CTAs) against a
  sequential reverse loop in fp32 torch ops on the device, and run-to-run
  determinism."""
  abi = _abi()
  torch.manual_seed(11)
  bsz, steps, width = 4, 2048, 2560
  a = (0.6 + 0.3999 * torch.rand(bsz, steps, width, device=DEV)).to(dtype)
  gy = torch.randn(bsz, steps, width, device=DEV).to(dtype)
  h = torch.randn(bsz, steps, width, device=DEV).to(dtype)
  reset = torch.rand(bsz, steps, device=DEV) < 0.002
  reset[:, 0] = True
  h0 = torch.randn(bsz, width, device=DEV)
  gl = torch.randn(bsz, width, device=DEV)
  dx, da, dh0 = abi.rnn_scan_bwd(gy, gl, a, h, reset, h0)
  dx2, da2, dh02 = abi.rnn_scan_bwd(gy, gl, a, h, reset, h0)
  assert torch.equal(dx, dx2) and torch.equal(da, da2) and torch.equal(dh0, dh02)
  am = (a * ~reset[..., None]).float()
  dh = gl.clone()
  want_dx = torch.empty(bsz, steps, width, device=DEV)
  for t in range(steps - 1, -1, -1):
    dh = dh + gy[:, t].float() if t == steps - 1 else am[:, t + 1] * dh + gy[:, t].float()
    want_dx[:, t] = dh
  hprev = torch.cat([h0[:, None], h[:, :-1].float()], dim=1)
  want_da = (want_dx * hprev) * ~reset[..., None]
  want_dh0 = am[:, 0] * want_dx[:, 0]
  if dtype == torch.float32:
    assert normwise(dx, want_dx) <= 1e-5 and normwise(da, want_da) <= 1e-5
  else:
    assert_close_bf16(dx.cpu(), want_dx.to(dtype).cpu(), "dx", min_identical=0.995)
    assert_close_bf16(da.cpu(), want_da.to(dtype).cpu(), "da", min_identical=0.995)
  assert normwise(dh0, want_dh0) <= 1e-5


@pytest.mark.parametrize("train_kernels", [True, False])
@pytest.mark.parametrize("case", fixture_io.cases("grad_rglru_"))
def test_rglru_training_path_golden(case, train_kernels):
  """RGLRU.forward with grad enabled -- gate math and scan forward / backward on
  the training kernels (train_kernels) or the reference's ATen op sequence around
  the differentiable scan: gradients of x, the cache and every parameter against
  the reference's autograd."""
  import cadence_gemma_b200 as cg
  from cadence_gemma_b200 import layers
  prev = layers.set_train_kernels(train_kernels)
  try:
    _rglru_training_golden(case)
  finally:
    layers.set_train_kernels(prev)


def _rglru_training_golden(case):
  import cadence_gemma_b200 as cg
  g = fixture_io.load(case)
  dtype = g["x"].dtype
  heads, bw, _ = g["param.input_gate.w"].shape
  lru = cg.RGLRU(heads * bw, heads, device=DEV, dtype=dtype)
  lru.load_state_dict({k[6:]: v for k, v in g.items() if k.startswith("param.")})
  x = cu(g["x"]).requires_grad_()
  cache = cu(g["cache"]).requires_grad_() if "cache" in g else None
  with torch.enable_grad():
    y, last_h = lru(x, cu(g["seg"]), cache)
    torch.autograd.backward([y, last_h], [cu(g["gy"]), cu(g["gh"])])
  bf = dtype == torch.bfloat16
  _close(y.detach(), g["y"], case + " y", min_identical=0.99)
  tol = 4e-2 if bf else 2e-5
  assert normwise(x.grad.float().cpu(), g["dx"].float()) <= tol, case
  if cache is not None:
    assert normwise(cache.grad.cpu(), g["dcache"]) <= (2e-2 if bf else 2e-5), case
  for name, prm in lru.named_parameters():
    nw = normwise(prm.grad.float().cpu(), g["grad." + name].float())
    assert nw <= (6e-2 if bf else 1e-4), (case, name, nw)
  with torch.no_grad():                       # the inference path is untouched by training mode
    y2, _ = lru(x.detach(), cu(g["seg"]), None if cache is None else cache.detach())
  _close(y2, g["y"], case + " y (no_grad)", min_identical=0.99)


@pytest.mark.parametrize("case", fixture_io.cases("grad_conv1d_"))
def test_conv1d_backward_golden(case):
  """cg_conv1d_bwd (dx, dw, db) against autograd through the reference Conv1D,
  fork mask included; second run bit-identical (fixed-order reductions)."""
  import cadence_gemma_b200 as cg
  g = fixture_io.load(case)
  dtype = g["x"].dtype
  conv = cg.Conv1D(g["x"].shape[-1], 4, device=DEV, dtype=dtype)
  conv.load_state_dict({"w": g["w"], "b": g["b"]})
  x = cu(g["x"]).requires_grad_()
  with torch.enable_grad():
    y, cache = conv(x, cu(g["seg"]))
    assert not cache.requires_grad
    y.backward(cu(g["gy"]))
  assert_bitexact(y.detach().cpu(), g["y"], case + " y")
  bf = dtype == torch.bfloat16
  assert_bitexact(x.grad.cpu(), g["dx"], case + " dx")        # autograd's bf16 / fp32 accumulation order
  for got, key in ((conv.w.grad, "dw"), (conv.b.grad, "db")):
    if bf:
      assert_close_bf16(got.cpu(), g[key], f"{case} {key}", min_identical=0.9)
    else:
      assert_close_f32(got.cpu(), g[key], f"{case} {key}", tol=2e-5)
  abi = _abi()
  r1 = abi.conv1d_bwd(cu(g["gy"]), x.detach(), conv.w.detach(), cu(g["seg"]))
  r2 = abi.conv1d_bwd(cu(g["gy"]), x.detach(), conv.w.detach(), cu(g["seg"]))
  assert all(torch.equal(u, v) for u, v in zip(r1, r2))


@pytest.mark.parametrize("mask_mode", [0, 1])
def test_conv1d_backward_full_size_vs_torch_autograd(mask_mode):
  """Config-2 sized Conv1D backward against torch autograd over an fp32
  restatement of the masked convolution on the device (both mask modes)."""
  abi = _abi()
  torch.manual_seed(21)
  bsz, steps, width = 4, 2048, 2560
  x = torch.randn(bsz, steps, width, device=DEV).bfloat16()
  gy = torch.randn(bsz, steps, width, device=DEV).bfloat16()
  w = (torch.randn(4, width, device=DEV) * 0.5).bfloat16()
  seg = torch.arange(steps, device=DEV, dtype=torch.int32)[None].repeat(bsz, 1)
  for b in range(bsz):
    cut = 100 + 333 * b
    seg[b, cut:] -= cut
    seg[b, cut + 1] = 0                           # adjacent document starts
  dx, dw, db = abi.conv1d_bwd(gy, x, w, seg, mask_mode=mask_mode)
  xf = x.float().requires_grad_()
  wf = w.float().requires_grad_()
  nz = (seg != 0)
  with torch.enable_grad():
    y = torch.zeros(bsz, steps, width, device=DEV)
    for s in range(4):
      xs = torch.nn.functional.pad(xf, (0, 0, s, 0))[:, :steps]
      m = torch.ones(bsz, steps, dtype=torch.bool, device=DEV)
      lo = 1 if mask_mode == 1 else 3              # fork: only tap 3 looks at seg[t-2]
      if s >= lo:
        for j in ((range(1, s + 1)) if mask_mode == 1 else [1]):
          # seg[t - s + j] != 0
          shifted = torch.nn.functional.pad(nz, (s - j, 0), value=True)[:, :steps]
          m = m & shifted
      y = y + xs * wf[3 - s] * m[..., None]
    y.backward(gy.float())
  assert_close_bf16(dx.cpu(), xf.grad.bfloat16().cpu(), "dx", min_identical=0.5)   # bf16 accumulation vs fp32
  assert normwise(dw.float(), wf.grad) <= 1e-2
  assert normwise(db.float(), gy.float().sum((0, 1))) <= 1e-2


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gate_kernels_vs_aten_autograd_full_size(dtype):
  """cg_rglru_gates_fwd / _bwd at a config-2 sized problem against the reference's
  ATen op sequence (layers.py:348-365, clipped sqrt gradient :224-238) evaluated
  under torch autograd on the device."""
  from cadence_gemma_b200 import layers
  abi = _abi()
  torch.manual_seed(31)
  bsz, steps, width = 4, 1024, 2560
  bf = dtype == torch.bfloat16
  x = torch.randn(bsz, steps, width, device=DEV).to(dtype).requires_grad_()
  px = (torch.randn(bsz, steps, width, device=DEV) * 1.5).to(dtype).requires_grad_()
  pa = (torch.randn(bsz, steps, width, device=DEV) * 1.5).to(dtype).requires_grad_()
  ap = (torch.rand(width, device=DEV) * -7 + 1).to(dtype).requires_grad_()
  reset = torch.rand(bsz, steps, device=DEV) < 0.01
  d_nx = torch.randn(bsz, steps, width, device=DEV).to(dtype)
  d_a = (torch.randn(bsz, steps, width, device=DEV) * ~reset[..., None]).to(dtype)
  with torch.enable_grad():
    gate_x, gate_a = torch.sigmoid(px), torch.sigmoid(pa)
    log_a = -8.0 * gate_a * torch.nn.functional.softplus(ap)
    a_ref = torch.exp(log_a)
    mult = layers.SqrtBoundDerivative.apply(1 - torch.exp(2 * log_a))
    mult = reset[..., None] + ~reset[..., None] * mult
    nx_ref = (x * gate_x) * mult.type(dtype)
    torch.autograd.backward([nx_ref, a_ref], [d_nx, d_a])
  a, nx = abi.rglru_gates_fwd(x.detach(), px.detach(), pa.detach(), ap.detach(), reset)
  if bf:   # same eager rounding points; libdevice expf vs ATen's: a few flips
    assert identical_fraction(a, a_ref.detach()) >= 0.999 and identical_fraction(nx, nx_ref.detach()) >= 0.995
  else:
    assert normwise(a, a_ref.detach()) <= 1e-5 and normwise(nx, nx_ref.detach()) <= 1e-5
  dx, dpx, dpa, dap = abi.rglru_gates_bwd(x.detach(), px.detach(), pa.detach(), ap.detach(), reset, d_nx, d_a)
  # bf16: the forward's saved tensors carry the eager rounding (the kernel reproduces
  # it), the chain rule itself runs in fp32 in the kernel and in bf16 in ATen
  tol = 3e-2 if bf else 2e-5
  for got, want, name in ((dx, x.grad, "dx"), (dpx, px.grad, "dpx"), (dpa, pa.grad, "dpa")):
    assert normwise(got.float(), want.float()) <= tol, (name, normwise(got.float(), want.float()))
  assert normwise(dap.float(), ap.grad.float()) <= (5e-2 if bf else 1e-4)
  r2 = abi.rglru_gates_bwd(x.detach(), px.detach(), pa.detach(), ap.detach(), reset, d_nx, d_a)
  assert all(torch.equal(u, v) for u, v in zip((dx, dpx, dpa, dap), r2))


@pytest.mark.parametrize("case", fixture_io.cases("grad_recurrent_block_"))
def test_recurrent_block_training_golden(case):
  """One training step's gradients through the RecurrentBlock mirror (linear_x /
  linear_y / linear_out by cuBLAS autograd, Conv1D and the scan by our backward
  kernels) against the reference block's autograd."""
  from cadence_gemma_b200.modules import RecurrentBlock
  g = fixture_io.load(case)
  dtype = g["x"].dtype
  blk = RecurrentBlock(width=64, num_heads=2, lru_width=256, conv1d_temporal_width=4,
                       device=DEV, dtype=dtype)
  blk.load_state_dict({k[6:]: v for k, v in g.items() if k.startswith("param.")})
  x = cu(g["x"]).requires_grad_()
  with torch.enable_grad():
    y, _ = blk(x, cu(g["seg"]))
    y.backward(cu(g["gy"]))
  bf = dtype == torch.bfloat16
  assert normwise(y.detach().float().cpu(), g["y"].float()) <= (2e-2 if bf else 2e-5)
  assert normwise(x.grad.float().cpu(), g["dx"].float()) <= (5e-2 if bf else 5e-5)
  for name, prm in blk.named_parameters():
    nw = normwise(prm.grad.float().cpu(), g["grad." + name].float())
    assert nw <= (8e-2 if bf else 2e-4), (case, name, nw)


# -------------------------------------------------------------------- conv1d
@pytest.mark.parametrize("case", fixture_io.cases("conv1d_"))
def test_conv1d_golden(case):
  abi = _abi()
  g = fixture_io.load(case)
  w, b = cu(g["w"]), cu(g["b"])
  if "x" in g:
    x = cu(g["x"])
    x_before = x.clone()
    y, cache = abi.conv1d_fwd(x, w, b, cu(g["seg"]))
    assert torch.equal(x, x_before), "input must stay untouched"
    assert_bitexact(y.cpu(), g["y"], case + " y")
    assert_bitexact(cache.cpu(), g["cache"], case + " cache")
  else:
    cache = cu(g["cache_in"])
  for i in range(2):
    if f"step{i}_x" not in g:
      break
    ys, cache = abi.conv1d_decode(cu(g[f"step{i}_x"]), w, b, cache)
    assert_bitexact(ys.cpu(), g[f"step{i}_y"], f"{case} step{i} y")
    assert_bitexact(cache.cpu(), g[f"step{i}_cache"], f"{case} step{i} cache")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("mask_mode", [0, 1])
def test_conv1d_random_vs_oracle(dtype, mask_mode):
  from oracle import c_oracle
  abi = _abi()
  g = torch.Generator().manual_seed(11 + mask_mode)
  bsz, steps, width = 3, 157, 192
  x = torch.randn((bsz, steps, width), generator=g).to(dtype)
  w = (torch.randn((4, width), generator=g) * 0.5).to(dtype)
  b = (torch.randn((width,), generator=g) * 0.1).to(dtype)
  seg = (torch.rand((bsz, steps), generator=g) > 0.2).long() * torch.arange(steps)[None]
  seg[:, :4] = -1
  y_ref, c_ref = c_oracle.conv1d_forward(w, b, x, seg, mask_mode=mask_mode)
  y, c = abi.conv1d_fwd(cu(x), cu(w), cu(b), cu(seg.to(torch.int32)),
                        mask_mode=mask_mode)
  assert_bitexact(y.cpu(), y_ref, "conv y")
  assert_bitexact(c.cpu(), c_ref, "conv cache")
  if dtype == torch.bfloat16:   # fp32-accumulate mode: within 2 bf16 ulp
    y32, _ = abi.conv1d_fwd(cu(x), cu(w), cu(b), cu(seg), mask_mode=mask_mode,
                            arith_mode=FP32)
    y_o, _ = c_oracle.conv1d_forward(w, b, x, seg, mask_mode=mask_mode,
                                     arith_mode=1)
    assert identical_fraction(y32.cpu(), y_o) >= 0.999


# --------------------------------------------------------------------- rglru
@pytest.mark.parametrize("case", fixture_io.cases("rglru_"))
@pytest.mark.parametrize("mode", [REF, REF | FAST, REF | STRICT, FP32, FP32 | FAST])
def test_rglru_golden_from_preacts(case, mode):
  """Kernel boundary: same pre-activations as the reference computed."""
  abi = _abi()
  g = fixture_io.load(case)
  bf = g["x"].dtype == torch.bfloat16
  x, px, pa = cu(g["x"]), cu(g["pre_x"]), cu(g["pre_a"])
  y, h = abi.rglru_fwd(x, px, pa, None, None, cu(g["a_param"]), cu(g["seg"]),
                       arith_mode=mode)
  if bf and (mode & FP32):
    # fp32-in-register mode is judged against the fp32-arithmetic oracle on
    # the same bf16 inputs (it does not reproduce the eager rounding points)
    from oracle import c_oracle
    y_ref, h_ref = c_oracle.rglru_from_preacts(g["x"], g["pre_x"], g["pre_a"],
                                               g["a_param"], g["seg"], arith_mode=1)
    _close(y, y_ref, f"{case} mode {mode} y", min_identical=0.99)
    assert normwise(h.cpu(), h_ref) <= 1e-4
    return
  _close(y, g["y"], f"{case} mode {mode} y")
  if bf:
    assert normwise(h.cpu(), g["last_h"]) <= 2e-2
  else:
    assert_close_f32(h.cpu(), g["last_h"], f"{case} mode {mode} h")
  cache = h
  for i in range(2):
    ys, cache = abi.rglru_fwd(cu(g[f"step{i}_x"]), cu(g[f"step{i}_pre_x"]),
                              cu(g[f"step{i}_pre_a"]), None, None, cu(g["a_param"]),
                              cu(g["seg"][:, -1:] + 1 + i), h0=cache,
                              arith_mode=mode)
    _close(ys, g[f"step{i}_y"], f"{case} mode {mode} step{i}", min_identical=0.9)


@pytest.mark.parametrize("case", fixture_io.cases("rglru_"))
def test_rglru_module_golden(case):
  """Module API incl. the cuBLAS gate GEMMs (state-dict keys as the reference)."""
  import cadence_gemma_b200 as cg
  g = fixture_io.load(case)
  width, heads = g["x"].shape[-1], g["input_gate_w"].shape[0]
  lru = cg.RGLRU(width, heads, device=DEV, dtype=g["x"].dtype)
  lru.load_state_dict({
      "a_param": g["a_param"], "input_gate.w": g["input_gate_w"],
      "input_gate.b": g["input_gate_b"], "a_gate.w": g["a_gate_w"],
      "a_gate.b": g["a_gate_b"]})
  with torch.no_grad():
    y, h = lru(cu(g["x"]), cu(g["seg"]))
    assert y.dtype == g["x"].dtype and h.dtype == torch.float32
    if y.dtype == torch.bfloat16:
      assert_close_bf16(y.cpu(), g["y"], case, min_identical=0.9)
    else:
      assert_close_f32(y.cpu(), g["y"], case, tol=2e-5)
    cache = h
    for i in range(2):
      ys, cache = lru(cu(g[f"step{i}_x"]), cu(g["seg"][:, -1:] + 1 + i), cache)
      if ys.dtype == torch.bfloat16:
        assert_close_bf16(ys.cpu(), g[f"step{i}_y"], f"{case} step{i}",
                          min_identical=0.8)
      else:
        assert_close_f32(ys.cpu(), g[f"step{i}_y"], f"{case} step{i}", tol=2e-5)
    y2, none = lru(cu(g["x"]), cu(g["seg"]), return_cache=False)
    assert none is None and torch.equal(y2, y)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 301, 256), (1, 1000, 64), (5, 33, 2560),
                                   (3, 1, 128)])
def test_rglru_random_vs_oracle(dtype, shape):
  """Ragged T / narrow E / bias split / h0 / int32 positions vs the C oracle."""
  from oracle import c_oracle, torch_port
  abi = _abi()
  bsz, steps, width = shape
  g = torch.Generator().manual_seed(steps * 7 + width)
  p = torch_port.init_rglru_params(width, max(1, width // 64), g, dtype)
  x = torch.randn(shape, generator=g).to(dtype)
  gx = torch.randn(shape, generator=g).to(dtype)
  ga = torch.randn(shape, generator=g).to(dtype)
  bx = p.input_gate_b.reshape(-1)
  ba = p.a_gate_b.reshape(-1)
  seg = torch.arange(steps)[None].repeat(bsz, 1)
  for b in range(bsz):
    cut = int(torch.randint(1, max(2, steps), (1,), generator=g))
    seg[b, cut:] -= cut
  h0 = torch.randn((bsz, width), generator=g)
  for mode in (REF, REF | FAST):
    y_ref, h_ref = c_oracle.rglru_from_preacts(x, gx, ga, p.a_param, seg, h0,
                                               bias_x=bx, bias_a=ba)
    y, h = abi.rglru_fwd(cu(x), cu(gx), cu(ga), cu(bx), cu(ba), cu(p.a_param),
                         cu(seg.to(torch.int32)), h0=cu(h0), arith_mode=mode)
    _close(y, y_ref, f"{shape} {dtype} mode {mode}", min_identical=0.97)
    assert normwise(h.cpu(), h_ref) <= (2e-2 if dtype == torch.bfloat16 else 1e-5)


# ------------------------------------------------------------ recurrent block
@pytest.mark.parametrize("case", fixture_io.cases("recurrent_block_"))
def test_recurrent_block_golden(case):
  from cadence_gemma_b200.modules import RecurrentBlock
  g = fixture_io.load(case)
  dtype = g["x"].dtype
  # shapes from the parameters (recurrent_block_bf16_h128 has head width 128: its
  # RG-LRU and gating product run on the fused tensor-core kernel)
  lru_width, width = g["param.linear_x.weight"].shape
  heads = g["param.rg_lru.input_gate.w"].shape[0]
  blk = RecurrentBlock(width=width, num_heads=heads, lru_width=lru_width, device=DEV, dtype=dtype)
  blk.load_state_dict({k[len("param."):]: v for k, v in g.items()
                       if k.startswith("param.")})
  bf = dtype == torch.bfloat16
  assert blk.rg_lru.uses_fused_kernel(cu(g["conv_in"])) == (bf and lru_width // heads in (128, 256))

  def check(a, b, what):
    if bf:
      assert_close_bf16(a.cpu(), b, what, min_identical=0.6)
    else:
      assert_close_f32(a.cpu(), b, what, tol=5e-5)

  with torch.no_grad():
    y, cache = blk(cu(g["x"]), cu(g["seg"]))
    check(y, g["y"], case + " y")
    if blk.rg_lru.uses_fused_kernel(cu(g["conv_in"])):
      # the optional folding of the gating product into the fused kernel's store
      #  gives the same bits as the separate multiply
      import cadence_gemma_b200 as cg
      old = cg.set_fold_gate(True)
      try:
        y_f, cache_f = blk(cu(g["x"]), cu(g["seg"]))
      finally:
        cg.set_fold_gate(old)
      assert torch.equal(y_f, y) and torch.equal(cache_f.rg_lru_state, cache.rg_lru_state)
    assert cache.rg_lru_state.dtype == torch.float32
    assert normwise(cache.rg_lru_state.cpu(), g["rg_lru_state"]) <= (3e-2 if bf else 5e-5)
    check(cache.conv1d_state, g["conv1d_state"], case + " conv state")
    for i in range(2):
      ys, cache = blk(cu(g[f"step{i}_x"]), cu(g["seg"][:, -1:] + 1 + i), cache)
      check(ys, g[f"step{i}_y"], f"{case} step{i}")
    # hot path alone on the tensors captured inside the reference block
    xc, conv_state = blk.conv_1d(cu(g["conv_in"]), cu(g["seg"]))
    out, _ = blk.rg_lru(xc, cu(g["seg"]))
    check(out, g["rglru_out"], case + " captured hot path")
    assert_bitexact(conv_state.cpu(), g["conv1d_state"], case + " conv cache")


def test_griffin_tiny_hot_path():
  """BASELINE config 1: the hot path inside the real tiny Griffin (fp32)."""
  import cadence_gemma_b200 as cg
  g = fixture_io.load("griffin_tiny_f32_t128")
  for blk in (0, 1):
    P = lambda k: g[f"blk{blk}_param.{k}"]
    conv = cg.Conv1D(256, 4, device=DEV, dtype=torch.float32)
    conv.load_state_dict({"w": P("conv_1d.w"), "b": P("conv_1d.b")})
    lru = cg.RGLRU(256, 8, device=DEV, dtype=torch.float32)
    lru.load_state_dict({k[len("rg_lru."):]: g[f"blk{blk}_param.{k}"]
                         for k in ("rg_lru.a_param", "rg_lru.input_gate.w",
                                   "rg_lru.input_gate.b", "rg_lru.a_gate.w",
                                   "rg_lru.a_gate.b")})
    with torch.no_grad():
      xc, conv_state = conv(cu(g[f"blk{blk}_conv_in"]), cu(g["seg"]))
      y, h = lru(xc, cu(g["seg"]))
    assert_bitexact(conv_state.cpu(), g[f"blk{blk}_conv1d_state"], "conv state")
    assert_close_f32(y.cpu(), g[f"blk{blk}_rglru_out"], f"blk{blk} out", tol=2e-5)
    assert_close_f32(h.cpu(), g[f"blk{blk}_rg_lru_state"], f"blk{blk} h", tol=2e-5)


# -------------------------------------------- full-size properties (config 2/4)
def _full_inputs(bsz, steps, width, dtype, seed, resets):
  from oracle import torch_port
  g = torch.Generator().manual_seed(seed)
  p = torch_port.init_rglru_params(width, 10, g, dtype)
  x = torch.randn((bsz, steps, width), generator=g).to(dtype).to(DEV)
  gx = (torch.randn((bsz, steps, width), generator=g) * 1.5).to(dtype).to(DEV)
  ga = (torch.randn((bsz, steps, width), generator=g) * 1.5).to(dtype).to(DEV)
  seg = torch.arange(steps, dtype=torch.int32)[None].repeat(bsz, 1)
  for b in range(bsz):
    for cut in sorted(torch.randint(1, steps, (resets,), generator=g).tolist()):
      seg[b, cut:] = torch.arange(steps - cut, dtype=torch.int32)
  return p, x, gx, ga, seg.to(DEV)


@pytest.mark.parametrize("dtype,steps,bsz", [(torch.bfloat16, 2048, 8),
                                             (torch.float32, 2048, 4),
                                             (torch.bfloat16, 8192, 2)])
def test_full_size_fast_vs_strict(dtype, steps, bsz):
  """RecurrentGemma-2B shapes: chunked kernel vs the sequential device oracle."""
  abi = _abi()
  width = 2560
  p, x, gx, ga, seg = _full_inputs(bsz, steps, width, dtype, 1, resets=7)
  args = (x, gx, ga, cu(p.input_gate_b.reshape(-1)), cu(p.a_gate_b.reshape(-1)),
          cu(p.a_param), seg)
  y_s, h_s = abi.rglru_fwd(*args, arith_mode=REF | STRICT)
  for mode in (REF, REF | FAST):
    y, h = abi.rglru_fwd(*args, arith_mode=mode)
    y2, h2 = abi.rglru_fwd(*args, arith_mode=mode)
    assert torch.equal(y, y2) and torch.equal(h, h2), "run-to-run determinism"
    if dtype == torch.bfloat16:
      frac = identical_fraction(y, y_s)
      assert frac >= (0.999 if mode == REF else 0.97), (mode, frac)
      torch.testing.assert_close(y.float(), y_s.float(), rtol=1e-2, atol=3e-2)
    else:
      assert normwise(y, y_s) <= 1e-5, (mode, normwise(y, y_s))
    assert normwise(h, h_s) <= (1e-2 if dtype == torch.bfloat16 else 1e-5)


def test_full_size_continuation_and_reset_isolation():
  """Config 4 properties: prefill(T) == prefill(T1) ; prefill(T2, h0), and a
  reset makes the output independent of everything before it."""
  abi = _abi()
  bsz, steps, width, dtype = 2, 8192, 2560, torch.float32
  p, x, gx, ga, seg = _full_inputs(bsz, steps, width, dtype, 3, resets=7)
  consts = (None, None, cu(p.a_param))
  y, h = abi.rglru_fwd(x, gx, ga, *consts, seg, arith_mode=REF)
  cut = 3000
  y1, h1 = abi.rglru_fwd(x[:, :cut], gx[:, :cut], ga[:, :cut], *consts,
                         seg[:, :cut], arith_mode=REF)
  y2, h2 = abi.rglru_fwd(x[:, cut:], gx[:, cut:], ga[:, cut:], *consts,
                         seg[:, cut:], h0=h1, arith_mode=REF)
  assert normwise(torch.cat([y1, y2], 1), y) <= 1e-5
  assert normwise(h2, h) <= 1e-5
  # reset isolation: force a reset at `cut`, scramble the prefix
  seg2 = seg.clone()
  seg2[:, cut] = 0
  ya, _ = abi.rglru_fwd(x, gx, ga, *consts, seg2, arith_mode=REF)
  xb = x.clone()
  xb[:, :cut] = torch.randn_like(xb[:, :cut]) * 3
  yb, _ = abi.rglru_fwd(xb, gx, ga, *consts, seg2, arith_mode=REF)
  assert torch.equal(ya[:, cut:], yb[:, cut:])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("width,heads", [(256, 8), (128, 16), (2560, 10)])
def test_fused_gate_layout_equals_separate(dtype, width, heads):
  """One fused [N,H,2bw] gate GEMM and two separate GEMM outputs give the same bits."""
  import cadence_gemma_b200 as cg
  abi = _abi()
  torch.manual_seed(5)
  lru = cg.RGLRU(width, heads, device=DEV, dtype=dtype)
  with torch.no_grad():
    lru.input_gate.b.normal_()
    lru.a_gate.b.normal_()
    x = torch.randn(2, 75, width, device=DEV).to(dtype)
    seg = torch.arange(75, device=DEV)[None].repeat(2, 1)
    seg[1, 40:] -= 40
    old = cg.set_fused(False)     # this test is about the cuBLAS gate-GEMM layouts
    try:
      y_f, h_f = lru(x, seg)
    finally:
      cg.set_fused(old)
    y_s, h_s = abi.rglru_fwd(x, lru.input_gate.gemm(x), lru.a_gate.gemm(x), lru.input_gate.b,
                             lru.a_gate.b, lru.a_param, seg, arith_mode=cg.get_arith_mode())
    y_t, _ = abi.rglru_fwd(x, None, None, lru.input_gate.b, lru.a_gate.b, lru.a_param, seg,
                           arith_mode=STRICT, gemm_fused=lru.gate_gemm(x),
                           block_width=width // heads)
  assert torch.equal(y_f, y_s) and torch.equal(h_f, h_s)
  if dtype == torch.bfloat16:
    assert identical_fraction(y_f, y_t) >= 0.97
  else:
    assert normwise(y_f, y_t) <= 1e-5


def test_abi_argument_errors():
  abi = _abi()
  x = torch.zeros(1, 4, 12, device=DEV, dtype=torch.bfloat16)   # E % 8 != 0
  seg = torch.zeros(1, 4, dtype=torch.int32, device=DEV)
  ap = torch.zeros(12, device=DEV, dtype=torch.bfloat16)
  with pytest.raises(AssertionError):
    abi.rglru_fwd(x, x, x, None, None, ap, seg)
  with pytest.raises(AssertionError):   # decode wants exactly one token
    abi.conv1d_decode(torch.zeros(1, 2, 16, device=DEV), torch.zeros(4, 16, device=DEV),
                      torch.zeros(16, device=DEV), torch.zeros(1, 3, 16, device=DEV))


# ----------------------------------------- fused tensor-core RG-LRU (tcgen05)
def _fused_cases():
  out = []
  for case in fixture_io.cases("rglru_bf16"):
    g = fixture_io.load(case)
    width, heads = g["x"].shape[-1], g["input_gate_w"].shape[0]
    if width // heads in (128, 256) and g["x"].shape[1] > 1:
      out.append(case)
  return out


@pytest.mark.parametrize("mode", [REF, REF | FAST])
@pytest.mark.parametrize("case", _fused_cases())
def test_fused_rglru_golden(case, mode):
  """cg_rglru_fused_fwd (gate GEMMs on tcgen05 + gates + scan) against the
  reference's own outputs, incl. the GEMM: pre-activations must come out
  bit-identical to the reference's bf16 einsum in >= 99.9 % of the elements."""
  abi = _abi()
  g = fixture_io.load(case)
  heads = g["input_gate_w"].shape[0]
  assert abi.fused_supported(g["x"].shape[-1], heads, torch.bfloat16)
  wpack = abi.pack_gate_weights(cu(g["input_gate_w"]), cu(g["a_gate_w"]))
  ws = abi.fused_workspace(torch.device(DEV), *g["x"].shape)
  y, h, dbg = abi.rglru_fused_fwd(cu(g["x"]), wpack, cu(g["input_gate_b"]), cu(g["a_gate_b"]),
                                  cu(g["a_param"]), cu(g["seg"]), heads, arith_mode=mode,
                                  debug=True, workspace=ws)
  # the debug planes hold round_bf16(x @ w) WITHOUT bias and the transposed x
  bx = g["input_gate_b"].reshape(-1).float()
  ba = g["a_gate_b"].reshape(-1).float()
  pre_x = (dbg[0].cpu().float() + bx).to(torch.bfloat16)
  pre_a = (dbg[1].cpu().float() + ba).to(torch.bfloat16)
  assert identical_fraction(pre_x, g["pre_x"]) >= 0.999, case
  assert identical_fraction(pre_a, g["pre_a"]) >= 0.999, case
  assert_bitexact(dbg[2].cpu(), g["x"], case + " x through the identity MMA")
  assert_close_bf16(y.cpu(), g["y"], f"{case} mode {mode}",
                    min_identical=0.995 if mode == REF else 0.98)
  assert normwise(h.cpu(), g["last_h"]) <= 1e-2
  # no debug planes: same bits
  y2, h2 = abi.rglru_fused_fwd(cu(g["x"]), wpack, cu(g["input_gate_b"]), cu(g["a_gate_b"]),
                               cu(g["a_param"]), cu(g["seg"]), heads, arith_mode=mode,
                               workspace=ws)
  assert torch.equal(y2, y) and torch.equal(h2, h)


@pytest.mark.parametrize("shape", [(1, 64, 256, 1), (3, 200, 512, 2), (2, 33, 256, 2),
                                   (1, 1000, 2560, 10), (5, 97, 1024, 4), (2, 31, 512, 4)])
def test_fused_rglru_equals_unfused(shape):
  """Fused kernel vs cuBLAS gate GEMM + scan kernel on the same inputs: ragged T,
  both head widths, h0, random document starts (incl. t = 0 .. 31 boundaries)."""
  import cadence_gemma_b200 as cg
  abi = _abi()
  bsz, steps, width, heads = shape
  g = torch.Generator().manual_seed(sum(shape))
  lru = cg.RGLRU(width, heads, device=DEV, dtype=torch.bfloat16)
  bw = width // heads
  with torch.no_grad():
    lru.input_gate.w.copy_((torch.randn((heads, bw, bw), generator=g) * bw ** -0.5).to(torch.bfloat16))
    lru.a_gate.w.copy_((torch.randn((heads, bw, bw), generator=g) * bw ** -0.5).to(torch.bfloat16))
    lru.input_gate.b.copy_(torch.randn((heads, bw), generator=g).to(torch.bfloat16))
    lru.a_gate.b.copy_(torch.randn((heads, bw), generator=g).to(torch.bfloat16))
  x = torch.randn((bsz, steps, width), generator=g).to(torch.bfloat16).to(DEV)
  seg = torch.arange(steps, dtype=torch.int32)[None].repeat(bsz, 1)
  for b in range(bsz):
    for cut in torch.randint(1, steps, (3,), generator=g).tolist():
      seg[b, cut:] = torch.arange(steps - cut, dtype=torch.int32)
  seg = seg.to(DEV)
  h0 = torch.randn((bsz, width), generator=g).to(DEV)
  for mode in (REF, REF | FAST):
    old_mode = cg.set_arith_mode(mode)
    try:
      assert lru.uses_fused_kernel(x)
      with torch.no_grad():
        y_f, h_f = lru(x, seg, h0)
        old = cg.set_fused(False)
        try:
          assert not lru.uses_fused_kernel(x)
          y_u, h_u = lru(x, seg, h0)
        finally:
          cg.set_fused(old)
    finally:
      cg.set_arith_mode(old_mode)
    assert identical_fraction(y_f, y_u) >= 0.999, (shape, mode, identical_fraction(y_f, y_u))
    torch.testing.assert_close(y_f.float(), y_u.float(), rtol=1e-2, atol=3e-2)
    assert normwise(h_f, h_u) <= 1e-5, (shape, mode, normwise(h_f, h_u))


@pytest.mark.parametrize("steps,bsz", [(2048, 8), (8192, 2)])
def test_fused_rglru_full_size_properties(steps, bsz):
  """RecurrentGemma-2B shapes (configs 2 and 4): determinism, agreement with the
  sequential device oracle, chunked continuation and reset isolation."""
  import cadence_gemma_b200 as cg
  abi = _abi()
  width, heads = 2560, 10
  p, x, _, _, seg = _full_inputs(bsz, steps, width, torch.bfloat16, 11, resets=7)
  lru = cg.RGLRU(width, heads, device=DEV, dtype=torch.bfloat16)
  with torch.no_grad():
    lru.a_param.copy_(p.a_param)
    lru.input_gate.b.copy_(p.input_gate_b)
    lru.a_gate.b.copy_(p.a_gate_b)
    y, h = lru(x, seg)
    y2, h2 = lru(x, seg)
    assert torch.equal(y, y2) and torch.equal(h, h2), "run-to-run determinism"
    # sequential one-thread-per-channel kernel on the cuBLAS pre-activations
    y_s, h_s = abi.rglru_fwd(x, None, None, lru.input_gate.b, lru.a_gate.b, lru.a_param, seg,
                             arith_mode=cg.get_arith_mode() | STRICT,
                             gemm_fused=lru.gate_gemm(x), block_width=width // heads)
    assert identical_fraction(y, y_s) >= 0.97
    torch.testing.assert_close(y.float(), y_s.float(), rtol=1e-2, atol=3e-2)
    assert normwise(h, h_s) <= 1e-2
    # continuation: prefill(T) == prefill(T1) ; prefill(T2, h0 = last_h)
    cut = 1000
    y1, h1 = lru(x[:, :cut].contiguous(), seg[:, :cut].contiguous())
    y2, h2 = lru(x[:, cut:].contiguous(), seg[:, cut:].contiguous(), h1)
    ycat = torch.cat([y1, y2], 1)
    assert identical_fraction(ycat, y) >= 0.9999
    assert normwise(h2, h) <= 1e-5
    # reset isolation: a document start makes y independent of everything before
    seg2 = seg.clone()
    seg2[:, cut] = 0
    ya, _ = lru(x, seg2)
    xb = x.clone()
    xb[:, :cut] = torch.randn_like(xb[:, :cut]) * 3
    yb, _ = lru(xb, seg2)
    assert torch.equal(ya[:, cut:], yb[:, cut:])
  torch.cuda.synchronize()   # a fired watchdog traps: the synchronisation would raise


def test_fused_rglru_argument_errors():
  abi = _abi()
  assert not abi.fused_supported(256, 8, torch.bfloat16)       # head width 32
  assert not abi.fused_supported(2560, 10, torch.float32)      # fp32 takes the scan kernel
  x = torch.zeros(1, 8, 256, device=DEV, dtype=torch.bfloat16)
  w = torch.zeros(2, 128, 128, device=DEV, dtype=torch.bfloat16)
  wpack = abi.pack_gate_weights(w, w)
  seg = torch.zeros(1, 8, dtype=torch.int32, device=DEV)
  ap = torch.zeros(256, device=DEV, dtype=torch.bfloat16)
  with pytest.raises(AssertionError):     # fp32-in-registers mode is not a mode of the fused kernel
    abi.rglru_fused_fwd(x, wpack, None, None, ap, seg, 2, arith_mode=FP32)
  with pytest.raises(AssertionError):     # workspace too small
    abi.rglru_fused_fwd(x, wpack, None, None, ap, seg, 2,
                        workspace=torch.zeros(64, dtype=torch.uint8, device=DEV))


def test_host_prefill_pipeline_matches_direct_call():
  """The pipelined host-buffer entry point returns the bits of the direct call."""
  import cadence_gemma_b200 as cg
  from cadence_gemma_b200.hostio import HostPrefill
  bsz, steps, width, heads = 8, 300, 512, 2
  torch.manual_seed(3)
  conv = cg.Conv1D(width, 4, device=DEV, dtype=torch.bfloat16)
  lru = cg.RGLRU(width, heads, device=DEV, dtype=torch.bfloat16)
  with torch.no_grad():
    conv.b.normal_()
    lru.input_gate.b.normal_()
    lru.a_gate.b.normal_()
  x = torch.randn(bsz, steps, width).to(torch.bfloat16).pin_memory()
  seg = torch.arange(steps, dtype=torch.int32)[None].repeat(bsz, 1)
  seg[:, 111:] -= 111
  seg = seg.pin_memory()
  y = torch.empty_like(x).pin_memory()
  h = torch.empty(bsz, width).pin_memory()
  c = torch.empty(bsz, 3, width, dtype=torch.bfloat16).pin_memory()
  for chunks in (1, 4, 8):
    y.zero_(); h.zero_(); c.zero_()
    HostPrefill(conv, lru, bsz, steps, chunks=chunks)(x, seg, y, h, c)
    torch.cuda.synchronize()
    with torch.no_grad():
      xc, cs = conv(x.to(DEV), seg.to(DEV))
      y_d, h_d = lru(xc, seg.to(DEV))
    assert torch.equal(y, y_d.cpu()) and torch.equal(h, h_d.cpu()) and torch.equal(c, cs.cpu()), chunks


def test_host_prefill_streaming_submit_overlaps_batches_correctly():
  """submit(): consecutive batches share the device buffers and overlap on the
  three streams; every batch must still return the bits of the direct call, and
  a blocking __call__ after submitted batches must drain them first."""
  import cadence_gemma_b200 as cg
  from cadence_gemma_b200.hostio import HostPrefill
  bsz, steps, width, heads = 4, 260, 512, 2
  torch.manual_seed(5)
  conv = cg.Conv1D(width, 4, device=DEV, dtype=torch.bfloat16)
  lru = cg.RGLRU(width, heads, device=DEV, dtype=torch.bfloat16)
  with torch.no_grad():
    conv.b.normal_()
    lru.input_gate.b.normal_()
    lru.a_gate.b.normal_()
  seg = torch.arange(steps, dtype=torch.int32)[None].repeat(bsz, 1)
  seg[:, 77:] -= 77
  seg = seg.pin_memory()
  nb = 5
  xs = [torch.randn(bsz, steps, width).to(torch.bfloat16).pin_memory() for _ in range(nb)]
  ys = [torch.zeros_like(xs[0]).pin_memory() for _ in range(nb)]
  hs = [torch.zeros(bsz, width).pin_memory() for _ in range(nb)]
  cs = [torch.zeros(bsz, 3, width, dtype=torch.bfloat16).pin_memory() for _ in range(nb)]
  for chunks in (1, 2, 4):
    hp = HostPrefill(conv, lru, bsz, steps, chunks=chunks)
    for t in ys + hs + cs:
      t.zero_()
    events = [hp.submit(xs[i], seg, ys[i], hs[i], cs[i]) for i in range(nb - 1)]
    hp(xs[-1], seg, ys[-1], hs[-1], cs[-1])              # blocking form after streamed ones
    for e in events:
      e.synchronize()
    torch.cuda.synchronize()
    for i in range(nb):
      with torch.no_grad():
        xc, cd = conv(xs[i].to(DEV), seg.to(DEV))
        y_d, h_d = lru(xc, seg.to(DEV))
      assert torch.equal(ys[i], y_d.cpu()) and torch.equal(hs[i], h_d.cpu()) and torch.equal(cs[i], cd.cpu()), (chunks, i)


def test_back_to_back_steps_with_programmatic_dependent_launch_are_ordered():
  """Conv1D -> prologue -> fused RG-LRU kernel are chained by programmatic
  dependent launches (the consumers start under their producer's tail).  Many
  unsynchronised steps on alternating inputs, with buffers reused across steps,
  must give exactly the results of fully synchronised single steps."""
  import cadence_gemma_b200 as cg
  torch.manual_seed(17)
  bsz, steps, width, heads = 8, 2048, 2560, 10
  conv = cg.Conv1D(width, 4, device=DEV, dtype=torch.bfloat16)
  lru = cg.RGLRU(width, heads, device=DEV, dtype=torch.bfloat16)
  with torch.no_grad():
    conv.w.normal_(0, 0.4)
    conv.b.normal_(0, 0.2)
    lru.input_gate.b.normal_()
    lru.a_gate.b.normal_()
  seg = torch.arange(steps, dtype=torch.int32, device=DEV)[None].repeat(bsz, 1)
  seg[:, 777:] -= 777
  xs = [torch.randn(bsz, steps, width, device=DEV).bfloat16() for _ in range(3)]
  want = []
  for x in xs:                                   # synchronised reference runs
    xc, cs = conv(x, seg)
    torch.cuda.synchronize()
    y, h = lru(xc, seg)
    torch.cuda.synchronize()
    want.append((y.clone(), h.clone(), cs.clone()))
  xc_buf = torch.empty_like(xs[0])
  y_bufs = [torch.empty_like(xs[0]) for _ in range(3)]
  h_bufs = [torch.empty(bsz, width, device=DEV) for _ in range(3)]
  c_bufs = [torch.empty(bsz, 3, width, device=DEV, dtype=torch.bfloat16) for _ in range(3)]
  for it in range(30):                           # no host sync; xc_buf is reused every step
    k = it % 3
    conv.forward_into(xs[k], seg, out=xc_buf, cache_out=c_bufs[k])
    lru.forward_into(xc_buf, seg, out=y_bufs[k], last_h_out=h_bufs[k])
    if it % 7 == 3:                              # an unrelated kernel in between now and then
      xc_buf.mul_(1.0)
  torch.cuda.synchronize()
  for k in range(3):
    assert torch.equal(y_bufs[k], want[k][0]) and torch.equal(h_bufs[k], want[k][1]), k
    assert torch.equal(c_bufs[k], want[k][2]), k
  abi = _abi()
  torch.cuda.synchronize()   # a fired watchdog traps: the synchronisation would raise


def test_fused_rglru_api_variants_and_cuda_graph():
  """Module API corner cases on the fused path: 1-D / int64 / broadcast
  segment_pos, no h0, return_cache=False, B = 1, and CUDA-graph capture of the
  whole call (prologue + tcgen05 kernel; the tensor map travels as a kernel
  parameter)."""
  import cadence_gemma_b200 as cg
  abi = _abi()
  torch.manual_seed(9)
  width, heads, steps = 512, 2, 130
  lru = cg.RGLRU(width, heads, device=DEV, dtype=torch.bfloat16)
  with torch.no_grad():
    lru.input_gate.b.normal_()
    lru.a_gate.b.normal_()
    x = torch.randn(3, steps, width, device=DEV).to(torch.bfloat16)
    seg = torch.cat([torch.arange(50), torch.arange(steps - 50)]).to(DEV)      # 1-D, int64
    assert lru.uses_fused_kernel(x)
    y1, h1 = lru(x, seg[None].repeat(3, 1))                                    # [B,T] int64
    y2, h2 = lru(x, seg[None].repeat(3, 1).to(torch.int32))
    assert torch.equal(y1, y2) and torch.equal(h1, h2)
    y3, none = lru(x, seg[None].repeat(3, 1), return_cache=False)
    assert none is None and torch.equal(y3, y1)
    yb, hb = lru(x[1:2].contiguous(), seg)                                     # B = 1, 1-D positions (:342-344)
    assert torch.equal(yb, y1[1:2]) and torch.equal(hb, h1[1:2])
    with pytest.raises(AssertionError):                                        # 1-D positions need B = 1, as the reference
      lru(x, seg)
    # CUDA graph: capture once, replay on new inputs in the static buffers
    xs, ys, hs = x.clone(), torch.empty_like(x), torch.empty_like(h1)
    seg32 = seg[None].repeat(3, 1).to(torch.int32)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
      lru.forward_into(xs, seg32, out=ys, last_h_out=hs)                       # warm (workspace, weights)
      torch.cuda.synchronize()
      g = torch.cuda.CUDAGraph()
      with torch.cuda.graph(g, stream=side):
        lru.forward_into(xs, seg32, out=ys, last_h_out=hs)
    torch.cuda.current_stream().wait_stream(side)
    for k in range(3):
      xn = torch.randn_like(x)
      xs.copy_(xn)
      g.replay()
      torch.cuda.synchronize()
      y_ref, h_ref = lru(xn, seg32)
      assert torch.equal(ys, y_ref) and torch.equal(hs, h_ref), k
  torch.cuda.synchronize()   # a fired watchdog traps: the synchronisation would raise


@pytest.mark.parametrize("shape", [(2, 100, 512, 2), (3, 64, 256, 2), (1, 333, 2560, 10)])
def test_fused_gating_product(shape):
  """`x * y` of RecurrentBlock.forward (modules.py:651) folded
  into the fused kernel's store == the separate bf16 multiply, bit for bit."""
  import cadence_gemma_b200 as cg
  bsz, steps, width, heads = shape
  torch.manual_seed(sum(shape))
  lru = cg.RGLRU(width, heads, device=DEV, dtype=torch.bfloat16)
  with torch.no_grad():
    lru.input_gate.b.normal_()
    lru.a_gate.b.normal_()
    x = torch.randn(bsz, steps, width, device=DEV).to(torch.bfloat16)
    gate = torch.randn(bsz, steps, width, device=DEV).to(torch.bfloat16)
    seg = torch.arange(steps, device=DEV, dtype=torch.int32)[None].repeat(bsz, 1)
    seg[:, steps // 3:] -= steps // 3
    h0 = torch.randn(bsz, width, device=DEV)
    y, h = lru(x, seg, h0)
    ym, hm = lru.forward_into(x, seg, h0, gate_mul=gate)
    assert torch.equal(ym, y * gate)
    assert torch.equal(hm, h)


def _conv_lru(width, heads, seed):
  import cadence_gemma_b200 as cg
  torch.manual_seed(seed)
  conv = cg.Conv1D(width, 4, device=DEV, dtype=torch.bfloat16)
  lru = cg.RGLRU(width, heads, device=DEV, dtype=torch.bfloat16)
  with torch.no_grad():
    conv.w.normal_(0, 0.4)
    conv.b.normal_(0, 0.2)
    lru.input_gate.b.normal_()
    lru.a_gate.b.normal_()
  return conv, lru


@pytest.mark.parametrize("mask_mode", [0, 1])
@pytest.mark.parametrize("shape", [(8, 2048, 2560, 10), (3, 200, 512, 2), (2, 33, 256, 2), (1, 1000, 2560, 10),
                                   (5, 97, 1024, 4), (2, 3, 256, 2), (1, 2, 512, 2), (4, 64, 256, 1),
                                   (2, 100, 4096, 16), (4, 3000, 2560, 10)])
def test_fused_conv_prefill_equals_two_kernel_path(shape, mask_mode):
  """cg_recurrent_prefill_fwd -- the temporal convolution INSIDE the fused tcgen05
  RG-LRU kernel (ONE launch) -- against the Conv1D kernel followed by the fused
  RG-LRU kernel: y, last_h and the returned conv cache bit for bit, the in-kernel
  conv output (debug tap) bit-exact with cg_conv1d_fwd (itself bit-exact with the
  reference, test_conv1d_golden); both document masks, ragged T, T < 4, head
  widths 128 / 256, document starts at tile boundaries, h0, gate product."""
  import cadence_gemma_b200 as cg
  from cadence_gemma_b200 import pipeline
  abi = _abi()
  bsz, steps, width, heads = shape
  conv, lru = _conv_lru(width, heads, sum(shape))
  conv.mask_mode = mask_mode
  with torch.no_grad():
    x = torch.randn(bsz, steps, width, device=DEV).to(torch.bfloat16)
    seg = torch.arange(steps, device=DEV, dtype=torch.int32)[None].repeat(bsz, 1)
    g = torch.Generator().manual_seed(sum(shape) + mask_mode)
    for b in range(bsz):                      # document starts, incl. next to 32-step tile edges
      cuts = torch.randint(1, max(steps, 2), (3,), generator=g).tolist() + [31, 32, 33, 34, 64]
      for cut in sorted(c for c in cuts[:3 + 2 * (b % 3)] if c < steps):
        seg[b, cut:] = torch.arange(steps - cut, dtype=torch.int32, device=DEV)
    h0 = torch.randn(bsz, width, device=DEV)
    old = pipeline.set_fused_conv(False)
    try:
      assert not pipeline.can_fuse_conv(conv, lru, x)
      xc_ref = torch.empty_like(x)
      y_ref, cs_ref, h_ref = cg.recurrent_hot_path(conv, lru, x, seg, lru_cache=h0, conv_out=xc_ref)
    finally:
      pipeline.set_fused_conv(old)
    pipeline.set_fused_conv(True)
    assert pipeline.can_fuse_conv(conv, lru, x)
    for it in range(3):
      y, cs, h = cg.recurrent_hot_path(conv, lru, x, seg, lru_cache=h0)
      assert torch.equal(cs, cs_ref), (shape, it)
      assert torch.equal(y, y_ref) and torch.equal(h, h_ref), (shape, it, identical_fraction(y, y_ref))
    # the conv output as the epilogue saw it (identity-MMA transpose of the in-place conv)
    _, _, _, dbg = abi.recurrent_prefill_fwd(x, conv.w, conv.b, cg.layers.packed_gate_weight(lru),
                                             lru.input_gate.b, lru.a_gate.b, lru.a_param, seg, heads,
                                             h0=h0, mask_mode=mask_mode, arith_mode=cg.get_arith_mode(),
                                             debug=True)
    assert_bitexact(dbg[2].cpu(), xc_ref.cpu(), f"{shape} in-kernel conv output")
    # no caches requested / gating product folded in
    y_n, cs_n, h_n = cg.recurrent_hot_path(conv, lru, x, seg, lru_cache=h0, return_cache=False)
    assert cs_n is None and h_n is None and torch.equal(y_n, y_ref)
    gate = torch.randn_like(x)
    y_g, _, _ = cg.recurrent_hot_path(conv, lru, x, seg, lru_cache=h0, gate_mul=gate)
    assert torch.equal(y_g, y_ref * gate)
    pipeline.set_fused_conv(old)
    # "auto": one launch at every size (it is the faster route everywhere since round 2's cluster version)
    assert pipeline.can_fuse_conv(conv, lru, x)
  torch.cuda.synchronize()


@pytest.mark.parametrize("case", fixture_io.cases("recurrent_block_"))
def test_fused_conv_prefill_recurrent_block_golden(case):
  """The reference's own RecurrentBlock outputs (tests/golden) against the hot path
  with the convolution inside the fused kernel, where the fixture's shape takes it."""
  import cadence_gemma_b200 as cg
  from cadence_gemma_b200 import pipeline
  g = fixture_io.load(case)
  if g["x"].dtype != torch.bfloat16:
    pytest.skip("fp32 fixture: not on the fused path")
  width = g["conv_in"].shape[-1]
  heads = g["param.rg_lru.input_gate.w"].shape[0]
  if width // heads not in (128, 256):
    pytest.skip("head width not on the fused path")
  conv = cg.Conv1D(width, 4, device=DEV, dtype=torch.bfloat16)
  lru = cg.RGLRU(width, heads, device=DEV, dtype=torch.bfloat16)
  with torch.no_grad():
    conv.w.copy_(cu(g["param.conv_1d.w"])); conv.b.copy_(cu(g["param.conv_1d.b"]))
    lru.a_param.copy_(cu(g["param.rg_lru.a_param"]))
    lru.input_gate.w.copy_(cu(g["param.rg_lru.input_gate.w"])); lru.input_gate.b.copy_(cu(g["param.rg_lru.input_gate.b"]))
    lru.a_gate.w.copy_(cu(g["param.rg_lru.a_gate.w"])); lru.a_gate.b.copy_(cu(g["param.rg_lru.a_gate.b"]))
    assert pipeline.can_fuse_conv(conv, lru, cu(g["conv_in"]))
    y, cs, h = cg.recurrent_hot_path(conv, lru, cu(g["conv_in"]), cu(g["seg"]))
  assert_bitexact(cs.cpu(), g["conv1d_state"], case + " conv cache")
  assert normwise(h.cpu(), g["rg_lru_state"]) <= 1e-2
  assert_close_bf16(y.cpu(), g["rglru_out"], case + " rg_lru output", min_identical=0.98)


@pytest.mark.parametrize("shape", [(32, 2560, 10), (5, 512, 2), (8, 256, 4), (1, 256, 2), (3, 1024, 4)])
@pytest.mark.parametrize("cache_dtype", [torch.bfloat16, torch.float32])
def test_fused_decode_step_equals_three_kernel_path(shape, cache_dtype):
  """cg_recurrent_decode_step (conv step + gate GEMVs + gates + h = a*h0 + x~ in one
  launch) vs conv1d_decode -> cuBLAS GEMV -> gate step: conv cache bit-exact, y and
  last_h equal up to the GEMV summation order; document starts, odd batch sizes,
  fp32 conv cache, missing h0, folded gating product."""
  import cadence_gemma_b200 as cg
  from cadence_gemma_b200 import pipeline
  bsz, width, heads = shape
  torch.manual_seed(sum(shape))
  conv = cg.Conv1D(width, 4, device=DEV, dtype=torch.bfloat16)
  lru = cg.RGLRU(width, heads, device=DEV, dtype=torch.bfloat16)
  with torch.no_grad():
    conv.w.normal_(0, 0.4)
    conv.b.normal_(0, 0.2)
    lru.input_gate.b.normal_()
    lru.a_gate.b.normal_()
    x = torch.randn(bsz, 1, width, device=DEV).to(torch.bfloat16)
    cache = torch.randn(bsz, 3, width, device=DEV).to(cache_dtype)
    h0 = torch.randn(bsz, width, device=DEV)
    seg = torch.full((bsz, 1), 17, device=DEV, dtype=torch.int32)
    if bsz > 2:
      seg[1, 0] = 0                        # a document start at this decode step
    assert pipeline.can_fuse_decode(conv, lru, x, cache)
    for h_in in (h0, None):
      y, cs, h = cg.recurrent_hot_path(conv, lru, x, seg, conv_cache=cache, lru_cache=h_in)
      old = cg.set_fused(False)
      try:
        assert not pipeline.can_fuse_decode(conv, lru, x, cache)
        y_u, cs_u, h_u = cg.recurrent_hot_path(conv, lru, x, seg, conv_cache=cache, lru_cache=h_in)
      finally:
        cg.set_fused(old)
      assert cs.dtype == cache_dtype and torch.equal(cs, cs_u)
      assert identical_fraction(y, y_u) >= 0.995, identical_fraction(y, y_u)
      torch.testing.assert_close(y.float(), y_u.float(), rtol=1e-2, atol=3e-2)
      assert normwise(h, h_u) <= 1e-2
    gate = torch.randn_like(x)
    y_g, _, _ = cg.recurrent_hot_path(conv, lru, x, seg, conv_cache=cache, lru_cache=h0, gate_mul=gate)
    y_p, _, _ = cg.recurrent_hot_path(conv, lru, x, seg, conv_cache=cache, lru_cache=h0)
    assert torch.equal(y_g, y_p * gate)
    y_n, cs_n, h_n = cg.recurrent_hot_path(conv, lru, x, seg, conv_cache=cache, lru_cache=h0,
                                           return_cache=False)
    assert cs_n is None and h_n is None and torch.equal(y_n, y_p)


@pytest.mark.parametrize("ctas", [1, 3, 7])
def test_fused_rglru_family_loop_small_grid(ctas):
  """Fewer CTAs than column families (as on a device with few SMs): a CTA then
  walks several families in turn, reloading the gate weights -- same bits as the
  full grid."""
  abi = _abi()
  torch.manual_seed(ctas)
  bsz, steps, width, heads = 3, 150, 1024, 4          # 8 families of 128 channels
  wx = (torch.randn(heads, 256, 256, device=DEV) / 16).to(torch.bfloat16)
  wa = (torch.randn(heads, 256, 256, device=DEV) / 16).to(torch.bfloat16)
  bx = torch.randn(width, device=DEV).to(torch.bfloat16)
  ba = torch.randn(width, device=DEV).to(torch.bfloat16)
  ap = torch.randn(width, device=DEV).to(torch.bfloat16)
  x = torch.randn(bsz, steps, width, device=DEV).to(torch.bfloat16)
  seg = torch.arange(steps, device=DEV, dtype=torch.int32)[None].repeat(bsz, 1)
  seg[:, 70:] -= 70
  h0 = torch.randn(bsz, width, device=DEV)
  wpack = abi.pack_gate_weights(wx, wa)
  ws = abi.fused_workspace(torch.device(DEV), bsz, steps, width)
  y_ref, h_ref = abi.rglru_fused_fwd(x, wpack, bx, ba, ap, seg, heads, h0=h0, arith_mode=FAST, workspace=ws)
  y, h = abi.rglru_fused_fwd(x, wpack, bx, ba, ap, seg, heads, h0=h0, arith_mode=FAST | (ctas << 8),
                             workspace=ws)
  assert torch.equal(y, y_ref) and torch.equal(h, h_ref)


@pytest.mark.parametrize("ctas", [2, 4, 6])
@pytest.mark.parametrize("heads,width", [(4, 1024), (8, 1024)])
def test_fused_conv_prefill_small_grid(ctas, heads, width):
  """The one-launch kernel (convolution inside) with fewer clusters / CTAs than heads, as on a device with
  few SMs: a cluster then walks several heads in turn -- weight reload, convolution taps of the other heads
  from global memory, tile descriptors with changing families -- and gives the bits of the full grid."""
  abi = _abi()
  torch.manual_seed(ctas + heads)
  bsz, steps = 3, 150
  bw = width // heads
  wx = (torch.randn(heads, bw, bw, device=DEV) / bw ** 0.5).to(torch.bfloat16)
  wa = (torch.randn(heads, bw, bw, device=DEV) / bw ** 0.5).to(torch.bfloat16)
  bx = torch.randn(width, device=DEV).to(torch.bfloat16)
  ba = torch.randn(width, device=DEV).to(torch.bfloat16)
  ap = torch.randn(width, device=DEV).to(torch.bfloat16)
  cw = (torch.randn(4, width, device=DEV) * 0.4).to(torch.bfloat16)
  cb = (torch.randn(width, device=DEV) * 0.2).to(torch.bfloat16)
  x = torch.randn(bsz, steps, width, device=DEV).to(torch.bfloat16)
  seg = torch.arange(steps, device=DEV, dtype=torch.int32)[None].repeat(bsz, 1)
  seg[:, 70:] -= 70
  h0 = torch.randn(bsz, width, device=DEV)
  wpack = abi.pack_gate_weights(wx, wa)
  ws = abi.fused_workspace(torch.device(DEV), bsz, steps, width)
  ref = abi.recurrent_prefill_fwd(x, cw, cb, wpack, bx, ba, ap, seg, heads, h0=h0, arith_mode=FAST, workspace=ws)
  got = abi.recurrent_prefill_fwd(x, cw, cb, wpack, bx, ba, ap, seg, heads, h0=h0, arith_mode=FAST | (ctas << 8),
                                  workspace=ws)
  torch.cuda.synchronize()
  for a, b in zip(got, ref):
    assert torch.equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 2048, 2560, 10), (2, 100, 512, 2), (3, 64, 256, 4)])
def test_graphed_hot_path_equals_eager(shape):
  """GraphedHotPath (the sampler's B = 1 prefill is host-bound): the step captured
  once in a CUDA graph and replayed on new inputs gives the bits of the eager call, with and without a
  state to continue from; head width 256 (one launch), 256 (one launch, small), 64 (conv kernel + cuBLAS
  + scan kernel)."""
  import cadence_gemma_b200 as cg
  bsz, steps, width, heads = shape
  conv, lru = _conv_lru(width, heads, sum(shape))
  with torch.no_grad():
    g = cg.GraphedHotPath(conv, lru, bsz, steps, with_h0=True)
    for it in range(3):
      x = torch.randn(bsz, steps, width, device=DEV).to(torch.bfloat16)
      seg = torch.arange(steps, device=DEV, dtype=torch.int32)[None].repeat(bsz, 1)
      if it > 0:
        seg[:, steps // 2:] = torch.arange(steps - steps // 2, device=DEV, dtype=torch.int32)
      h0 = torch.randn(bsz, width, device=DEV) if it != 1 else None
      y_ref, cs_ref, h_ref = cg.recurrent_hot_path(conv, lru, x, seg, lru_cache=h0)
      y, cs, h = g(x, seg, h0)
      torch.cuda.synchronize()
      assert torch.equal(y, y_ref) and torch.equal(cs, cs_ref) and torch.equal(h, h_ref), (shape, it)
