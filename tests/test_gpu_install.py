"""The zero-edit route on hardware: the reference's OWN classes (imported unmodified
from baseline/_ref -- or /root/reference in the build container -- through
oracle/ref_loader.py) with ``install()`` patching the hot path in:

  * ``RecurrentBlock`` (reference modules.py:503-685) and ``ResidualBlock``
    (:688-914): reference eager CUDA run vs the same modules after ``install()``
    -- prefill + two cached decode steps, fused and unfused shapes, fp32 and bf16,
    also inside ``torch.utils.checkpoint(use_reentrant=False)`` as
    ``Griffin.forward`` calls its blocks (griffin.py:199-208);
  * tiny ``Griffin`` (BASELINE config 1: width 256, R,R,A, fp32): logits of the
    installed model vs the reference's eager run, vs the stored golden logits of
    ``griffin_tiny_f32_t128`` and at T = 512 (two B = 1 rows, quirk D4);
  * the ctypes stubs of INTEGRATION.md section 3, extracted from the markdown and
    EXECUTED against the library.

The reference run on CUDA is the tightest oracle tier of SURVEY.md 8(c) (same
libdevice math as the kernels), slow (a Python loop over T) but fine at test sizes.
"""
import contextlib
import io
import os
import re

import pytest
import torch

from tests.golden import fixture_io
from tests.helpers import identical_fraction, normwise

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref():
  from oracle import ref_loader
  if not ref_loader.reference_available():
    pytest.skip("no copy of the reference on this box (baseline/_ref missing: run __graft_entry__.build())")
  return ref_loader.load_reference_griffin()


@contextlib.contextmanager
def _installed(ref):
  from cadence_gemma_b200 import install
  install.install(ref.layers, ref.modules)
  try:
    yield
  finally:
    install.uninstall()


def _randomise(block, g):
  """Non-trivial gate / conv biases and conv taps (the reference zero-inits the biases)."""
  rb = block.recurrent_block if hasattr(block, "recurrent_block") else block
  with torch.no_grad():
    for p in (rb.rg_lru.input_gate.b, rb.rg_lru.a_gate.b, rb.conv_1d.b):
      p.copy_((torch.randn(p.shape, generator=g) * 0.5).to(p.dtype))
    rb.conv_1d.w.copy_((torch.randn(rb.conv_1d.w.shape, generator=g) * 0.4).to(rb.conv_1d.w.dtype))


def _seg(bsz, steps):
  seg = torch.arange(steps, dtype=torch.int32)[None].repeat(bsz, 1)
  seg[:, steps // 2:] = torch.arange(steps - steps // 2, dtype=torch.int32)   # test_utils.py:74-75 pattern
  return seg


def _run_block(block, x, seg, steps_x, use_checkpoint):
  """prefill + cached decode steps of a (Recurrent|Residual)Block; returns all outputs."""
  from torch.utils import checkpoint
  outs = []
  with torch.no_grad():
    if use_checkpoint:   # as Griffin.forward does (griffin.py:199-208)
      y, cache = checkpoint.checkpoint(block, x, seg, None, True, use_reentrant=False,
                                       determinism_check="none")
    else:
      y, cache = block(x, seg, None, True)
    outs.append(y)
    pos = seg[:, -1:].clone()
    for xs in steps_x:
      pos = pos + 1
      ys, cache = block(xs, pos, cache, True)
      outs.append(ys)
  inner = cache
  return outs, inner


def _compare(got, want, dtype, what, floor):
  if dtype == torch.bfloat16:
    torch.testing.assert_close(got.float(), want.float(), rtol=1e-2, atol=3e-2, msg=what)
    assert identical_fraction(got, want) >= floor, (what, identical_fraction(got, want))
  else:
    assert normwise(got, want) <= 5e-5, (what, normwise(got, want))


@pytest.mark.parametrize("use_checkpoint", [False, True])
@pytest.mark.parametrize("kind", ["recurrent", "residual"])
@pytest.mark.parametrize("dtype,width,lru_width,heads", [
    (torch.bfloat16, 256, 512, 2),     # head width 256: fused tcgen05 kernel, one-launch prefill
    (torch.bfloat16, 128, 256, 2),     # head width 128
    (torch.bfloat16, 128, 128, 4),     # head width 32: cuBLAS gate GEMM + scan kernel
    (torch.float32, 64, 128, 2),       # fp32
])
def test_reference_blocks_with_install(dtype, width, lru_width, heads, kind, use_checkpoint):
  ref = _ref()
  g = torch.Generator().manual_seed(width + lru_width + heads)
  torch.manual_seed(17)
  R = ref.common.TemporalBlockType.RECURRENT
  if kind == "recurrent":
    block = ref.modules.RecurrentBlock(width=width, num_heads=heads, lru_width=lru_width, device=DEV, dtype=dtype)
  else:
    block = ref.modules.ResidualBlock(width=width, mlp_expanded_width=3 * width, num_heads=heads,
                                      attention_window_size=2048, temporal_block_type=R,
                                      lru_width=lru_width, device=DEV, dtype=dtype)
  block.eval()
  _randomise(block, g)
  bsz, steps = 3, 100
  x = (torch.randn((bsz, steps, width), generator=g) * 0.7).to(dtype).to(DEV)
  seg = _seg(bsz, steps).to(DEV)
  steps_x = [(torch.randn((bsz, 1, width), generator=g) * 0.7).to(dtype).to(DEV) for _ in range(2)]

  want, want_cache = _run_block(block, x, seg, steps_x, use_checkpoint)     # the reference, eager CUDA
  with _installed(ref):
    assert ref.modules.RecurrentBlock.forward.__module__ == "cadence_gemma_b200.install"
    got, got_cache = _run_block(block, x, seg, steps_x, use_checkpoint)
  torch.cuda.synchronize()
  assert ref.modules.RecurrentBlock.forward.__module__ != "cadence_gemma_b200.install"   # uninstalled

  # block outputs go through linear_out (and the MLP): compare within the reference's tolerance
  for i, (a, b) in enumerate(zip(got, want)):
    _compare(a, b, dtype, f"{kind} {dtype} output {i}", floor=0.90 if kind == "recurrent" else 0.80)
  gc = got_cache if kind == "recurrent" else got_cache
  wc = want_cache
  assert gc.rg_lru_state.dtype == torch.float32 and gc.conv1d_state.dtype == wc.conv1d_state.dtype
  assert torch.equal(gc.conv1d_state, wc.conv1d_state), "conv cache after two decode steps"
  assert normwise(gc.rg_lru_state, wc.rg_lru_state) <= (2e-2 if dtype == torch.bfloat16 else 5e-5)


def test_reference_recurrent_block_with_install_matches_golden():
  """The installed reference block against the fixtures the reference produced on the
  CPU (tests/golden/recurrent_block_*.npz): prefill + two cached steps."""
  ref = _ref()
  for case in fixture_io.cases("recurrent_block_"):
    g = fixture_io.load(case)
    dtype = g["x"].dtype
    lru_width, width = g["param.linear_x.weight"].shape
    heads = g["param.rg_lru.input_gate.w"].shape[0]
    block = ref.modules.RecurrentBlock(width=width, num_heads=heads, lru_width=lru_width, device=DEV, dtype=dtype)
    block.load_state_dict({k[len("param."):]: v for k, v in g.items() if k.startswith("param.")})
    block.eval()
    with _installed(ref), torch.no_grad():
      y, cache = block(g["x"].to(DEV), g["seg"].to(DEV), None, True)
      _compare(y.cpu(), g["y"], dtype, case + " y", floor=0.6)
      # rows of linear_x's output: cuBLAS here, MKL in the fixture (not bit-comparable)
      _compare(cache.conv1d_state.cpu(), g["conv1d_state"], dtype, case + " conv cache", floor=0.9)
      for i in range(2):
        ys, cache = block(g[f"step{i}_x"].to(DEV), (g["seg"][:, -1:] + 1 + i).to(DEV), cache, True)
        _compare(ys.cpu(), g[f"step{i}_y"], dtype, f"{case} step{i}", floor=0.6)
    torch.cuda.synchronize()


def _tiny_griffin(ref, gradient_checkpointing):
  common = ref.common
  R, A = common.TemporalBlockType.RECURRENT, common.TemporalBlockType.ATTENTION
  cfg = common.GriffinConfig(
      vocab_size=1000, width=256, mlp_expanded_width=768, num_heads=8,
      lru_width=256, block_types=(R, R, A), embeddings_scale_by_sqrt_dim=True,
      attention_window_size=2048, logits_soft_cap=30.0)
  torch.manual_seed(0)
  with contextlib.redirect_stdout(io.StringIO()):
    model = ref.griffin.Griffin(cfg, gradient_checkpointing=gradient_checkpointing, dtype=torch.float32)
  model.eval()
  g = torch.Generator().manual_seed(0)
  with torch.no_grad():   # exactly what tests/golden/make_golden.py::make_griffin_tiny does
    for blk in model.blocks:
      if hasattr(blk, "recurrent_block"):
        rb = blk.recurrent_block
        for p in (rb.rg_lru.input_gate.b, rb.rg_lru.a_gate.b, rb.conv_1d.b):
          p.copy_(torch.randn(p.shape, generator=g) * 0.5)
        rb.conv_1d.w.copy_(torch.randn(rb.conv_1d.w.shape, generator=g) * 0.4)
  return model, g


@pytest.mark.parametrize("gradient_checkpointing", [False, True])
def test_tiny_griffin_with_install_config1(gradient_checkpointing):
  """BASELINE config 1 through the reference's own ``Griffin.forward`` with the kernels
  installed: the stored golden logits (T = 128, CPU reference) and, at T = 512 with a
  reset at 256, two B = 1 rows (the fork's forward cannot take B > 1, quirk D4)
  against the reference's eager CUDA run."""
  ref = _ref()
  model, g = _tiny_griffin(ref, gradient_checkpointing)
  fx = fixture_io.load("griffin_tiny_f32_t128")
  model = model.to(DEV)
  with _installed(ref), torch.no_grad():
    logits, cache = model(fx["tokens"].to(DEV), fx["seg"].to(DEV), return_logits=True, return_cache=True)
  assert normwise(logits.cpu(), fx["logits"]) <= 2e-4, normwise(logits.cpu(), fx["logits"])
  for key, c in cache.items():
    if hasattr(c, "rg_lru_state"):
      idx = key.split(".")[-1]
      assert normwise(c.rg_lru_state.cpu(), fx[f"blk{idx}_rg_lru_state"]) <= 5e-5
      # the conv cache holds rows of linear_x's output: cuBLAS here, MKL in the fixture
      assert normwise(c.conv1d_state.cpu(), fx[f"blk{idx}_conv1d_state"]) <= 1e-5
  # config 1 proper: T = 512, two rows run one at a time
  steps = 512
  g2 = torch.Generator().manual_seed(5)
  for row in range(2):
    tokens = torch.randint(0, 1000, (1, steps), generator=g2).to(DEV)
    seg = torch.cat([torch.arange(256), torch.arange(256)])[None].to(torch.int32).to(DEV)
    with torch.no_grad():
      want, _ = model(tokens, seg, return_logits=True, return_cache=True)          # reference, eager CUDA
      with _installed(ref):
        got, got_cache = model(tokens, seg, return_logits=True, return_cache=True)
        # one cached decode step on top (prefill -> decode continuity through the model)
        nxt = torch.randint(0, 1000, (1, 1), generator=g2).to(DEV)
        pos = torch.full((1, 1), 256, dtype=torch.int32, device=DEV)
        got1, _ = model(nxt, pos, cache=got_cache, return_logits=True, return_cache=True)
      want_pre, want_cache = model(tokens, seg, return_logits=True, return_cache=True)
      want1, _ = model(nxt, pos, cache=want_cache, return_logits=True, return_cache=True)
    assert normwise(got, want) <= 2e-4, (row, normwise(got, want))
    assert normwise(got1, want1) <= 2e-4, (row, normwise(got1, want1))
  torch.cuda.synchronize()


def _md_blocks():
  with open(os.path.join(ROOT, "INTEGRATION.md")) as f:
    text = f.read()
  sec = text[text.index("## 3. In-tree route"):text.index("## 4. Behavioural contract")]
  return re.findall(r"```python\n(.*?)```", sec, flags=re.S)


def test_integration_md_stubs_run():
  """INTEGRATION.md section 3, extracted and executed: the three stubs a maintainer would
  paste into layers.py / modules.py give the same bits as the package's own binding."""
  import types
  import cadence_gemma_b200 as cg
  from cadence_gemma_b200 import _abi, pipeline
  blocks = _md_blocks()
  assert len(blocks) == 3
  ns = {}
  for code in blocks:
    exec(compile(code.replace('"libcadence_b200.so"', repr(_abi.LIB_PATH)), "INTEGRATION.md", "exec"), ns)  # pylint: disable=exec-used
  torch.manual_seed(3)
  width, heads, bsz, steps = 512, 2, 3, 150
  conv = cg.Conv1D(width, 4, device=DEV, dtype=torch.bfloat16)
  lru = cg.RGLRU(width, heads, device=DEV, dtype=torch.bfloat16)
  with torch.no_grad():
    conv.w.normal_(0, 0.4); conv.b.normal_(0, 0.2)
    lru.input_gate.b.normal_(); lru.a_gate.b.normal_()
    x = torch.randn(bsz, steps, width, device=DEV).to(torch.bfloat16)
    seg = _seg(bsz, steps).to(DEV)
    h0 = torch.randn(bsz, width, device=DEV)
    # stub 1: cuBLAS gate GEMM + cg_rglru_fwd  ==  the unfused route of the package
    old = cg.set_fused(False)
    try:
      y_u, h_u = lru(x, seg, h0)
    finally:
      cg.set_fused(old)
    y1, h1 = ns["_rglru_cuda"](lru, x, seg, h0)
    assert torch.equal(y1, y_u) and torch.equal(h1, h_u)
    # stub 2: cg_rglru_fused_fwd  ==  the fused route; repacks after a weight update
    y_f, h_f = lru(x, seg, h0)
    y2, h2 = ns["_rglru_fused"](lru, x, seg, h0)
    assert torch.equal(y2, y_f) and torch.equal(h2, h_f)
    lru.input_gate.w.mul_(0.5)
    y_f2, _ = lru(x, seg, h0)
    y2b, _ = ns["_rglru_fused"](lru, x, seg, h0)
    assert torch.equal(y2b, y_f2) and not torch.equal(y2b, y2)
    # stub 3: cg_recurrent_prefill_fwd inside RecurrentBlock.forward
    old = pipeline.set_fused_conv(True)
    try:
      y_p, cs_p, h_p = cg.recurrent_hot_path(conv, lru, x, seg)
    finally:
      pipeline.set_fused_conv(old)
    block = types.SimpleNamespace(conv_1d=conv, rg_lru=lru)
    y3, cs3, h3 = ns["_conv_rglru_prefill"](block, x, seg)
    assert torch.equal(y3, y_p) and torch.equal(cs3, cs_p) and torch.equal(h3, h_p)
  torch.cuda.synchronize()
