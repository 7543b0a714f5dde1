"""Empty batch (B = 0, T >= 1): the shims return the reference's empty tensors
without launching anything.

Expected shapes / dtypes were recorded from the unmodified reference (probe in
this container: `rnn_scan`, `RGLRU.forward`, `Conv1D.forward` of
recurrentgemma/torch/layers.py:146-199, :322-375, :458-546 with B = 0): `y`
`[0,T,E]` in x.dtype, `last_h` `[0,E]` fp32, conv cache `[0,W-1,E]` in x.dtype
(prefill) or the cache's dtype (decode, :542).  When the reference is importable
(this container only) the test also runs it side by side.  T = 0 is an error in
the reference (`x[:, 0]`, :178) and stays one here.
"""
import pytest
import torch

import cadence_gemma_b200 as cg
from oracle import ref_loader

E, H, W = 64, 2, 4


def _meta(t):
  return None if t is None else (tuple(t.shape), t.dtype)


def _devices():
  return ["cpu"] + (["cuda"] if torch.cuda.is_available() else [])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("steps", [1, 5])
def test_empty_batch_shapes_and_dtypes(dtype, steps):
  for dev in _devices():
    x = torch.zeros((0, steps, E), dtype=dtype, device=dev)
    seg = torch.zeros((0, steps), dtype=torch.int32, device=dev)
    reset = torch.zeros((0, steps), dtype=torch.bool, device=dev)
    for h0 in (None, torch.zeros((0, E), device=dev)):
      y, h = cg.rnn_scan(x, x, reset, h0)
      assert _meta(y) == ((0, steps, E), dtype) and _meta(h) == ((0, E), torch.float32)
    lru = cg.RGLRU(E, H, device=dev, dtype=dtype)
    y, h = lru(x, seg)
    assert _meta(y) == ((0, steps, E), dtype) and _meta(h) == ((0, E), torch.float32)
    y, h = lru(x, seg, None, False)
    assert _meta(y) == ((0, steps, E), dtype) and h is None
    conv = cg.Conv1D(E, W, device=dev, dtype=dtype)
    y, c = conv(x, seg)
    assert _meta(y) == ((0, steps, E), dtype) and _meta(c) == ((0, W - 1, E), dtype)
    y, c = conv(x, seg, return_cache=False)
    assert _meta(y) == ((0, steps, E), dtype) and c is None
    if steps == 1:                                  # decode step: the cache keeps ITS dtype
      y, c = conv(x, seg, torch.zeros((0, W - 1, E), dtype=torch.float32, device=dev))
      assert _meta(y) == ((0, 1, E), dtype) and _meta(c) == ((0, W - 1, E), torch.float32)


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present")
@pytest.mark.parametrize("steps", [1, 4])
def test_empty_batch_matches_reference(steps):
  layers = ref_loader.load_reference().layers
  x = torch.zeros((0, steps, E))
  seg = torch.zeros((0, steps), dtype=torch.int64)
  reset = torch.zeros((0, steps), dtype=torch.bool)
  want = [layers.rnn_scan(x, x, reset, None), layers.RGLRU(E, H)(x, seg), layers.Conv1D(E, W)(x, seg)]
  got = [cg.rnn_scan(x, x, reset, None), cg.RGLRU(E, H)(x, seg), cg.Conv1D(E, W)(x, seg)]
  for w, g in zip(want, got):
    assert [_meta(t) for t in w] == [_meta(t) for t in g]


def test_install_routes_empty_batches():
  """The monkey-patched forwards (install.py) take the same early exit."""
  from cadence_gemma_b200 import install
  x = torch.zeros((0, 3, E), dtype=torch.bfloat16)
  seg = torch.zeros((0, 3), dtype=torch.int32)
  y, h = install._rglru_forward(cg.RGLRU(E, H, dtype=torch.bfloat16), x, seg)
  assert _meta(y) == ((0, 3, E), torch.bfloat16) and _meta(h) == ((0, E), torch.float32)
  y, c = install._conv1d_forward(cg.Conv1D(E, W, dtype=torch.bfloat16), x, seg)
  assert _meta(y) == ((0, 3, E), torch.bfloat16) and _meta(c) == ((0, W - 1, E), torch.bfloat16)


@pytest.mark.gpu
def test_zero_steps_is_an_error_like_the_reference():
  x = torch.zeros((2, 0, E), device="cuda")
  seg = torch.zeros((2, 0), dtype=torch.int32, device="cuda")
  with pytest.raises((AssertionError, IndexError, RuntimeError)):
    cg.RGLRU(E, H, device="cuda")(x, seg)
    torch.cuda.synchronize()
