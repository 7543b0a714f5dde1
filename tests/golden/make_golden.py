"""Generates the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container (where ``/root/reference`` is mounted):

    cd /root/repo && python tests/golden/make_golden.py

The reference has no golden vectors / known-answer tests for the recurrent
path, so parity is pinned against outputs of
the reference code itself executed here on CPU.  The fixtures are small
(`.npz`, bf16 stored as raw uint16) and are what travels to the GPU box, where
``/root/reference`` does not exist.

Case shapes follow the reference's own test template
(``recurrentgemma/torch/test_utils.py:59-107``): one mid-sequence reset
(``segment_pos = concat(arange(L/2), arange(L/2))``) followed by two
single-token decode steps fed with the returned cache, plus the edge cases of
SURVEY.md section 8 (T < temporal width, 1-D segment_pos, -1 left padding,
fp32 conv cache under bf16 activations, T == 1 scan with/without h0).
"""
from __future__ import annotations

import os
import sys
import zlib

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from tests.golden import fixture_io  # noqa: E402

DTYPES = {"f32": torch.float32, "bf16": torch.bfloat16}
# MAKE_GOLDEN_ONLY=substr[,substr]: regenerate only the RG-LRU fixtures whose name matches
ONLY = [o for o in os.environ.get("MAKE_GOLDEN_ONLY", "").split(",") if o]


def gen(seed):
  return torch.Generator().manual_seed(seed)


def seed_of(*parts):
  """Deterministic seed (Python's hash() is salted per process)."""
  return zlib.crc32(repr(parts).encode()) % 2**31


def halves(length, batch):
  half = torch.arange(length // 2)
  row = torch.cat([half, half, torch.arange(length - 2 * (length // 2))])
  return row[None, :].repeat(batch, 1)


def ragged_segments(batch, length, seed, pad=0):
  """Random document boundaries (incl. adjacent ones) and `pad` -1 paddings."""
  g = gen(seed)
  seg = torch.zeros((batch, length), dtype=torch.int64)
  for b in range(batch):
    pos = 0
    for t in range(length):
      if t < pad:
        seg[b, t] = -1
        continue
      if t == pad or torch.rand((), generator=g).item() < 0.15:
        pos = 0
      seg[b, t] = pos
      pos += 1
  return seg


# ---------------------------------------------------------------------------
def make_rnn_scan(ref):
  for tag, dtype in DTYPES.items():
    for name, (bsz, steps, width, with_h0) in {
        "t37_h0": (2, 37, 48, True),
        "t37_noh0": (2, 37, 48, False),
        "t1_h0": (3, 1, 40, True),
        "t1_noh0": (3, 1, 40, False),
        "t130_h0": (1, 130, 72, True),
    }.items():
      g = gen(seed_of(tag, name))
      x = torch.randn((bsz, steps, width), generator=g).to(dtype)
      a = (0.5 + 0.4999 * torch.rand((bsz, steps, width), generator=g)).to(dtype)
      reset = torch.rand((bsz, steps), generator=g) < 0.1
      reset[0, 0] = True
      h0 = torch.randn((bsz, width), generator=g) if with_h0 else None
      y, h_last = ref.layers.rnn_scan(x.clone(), a.clone(), reset, h0)
      fixture_io.save(f"rnn_scan_{tag}_{name}", dict(
          x=x, a=a, reset=reset, h0=h0, y=y, h_last=h_last))


def make_conv1d(ref):
  for tag, dtype in DTYPES.items():
    for tw, steps, width, bsz, segkind in [
        (4, 1, 32, 2, "halves"), (4, 2, 32, 2, "halves"),
        (4, 3, 32, 2, "halves"), (4, 4, 32, 2, "halves"),
        (4, 5, 32, 2, "ragged"), (4, 32, 64, 2, "halves"),
        (4, 33, 64, 3, "ragged"), (4, 32, 64, 2, "ragged_pad"),
        (4, 16, 32, 2, "oned"), (8, 32, 32, 2, "halves"),
        (8, 33, 32, 2, "ragged"), (2, 9, 32, 1, "ragged"),
    ]:
      seed = seed_of(tag, tw, steps, segkind)
      g = gen(seed)
      torch.manual_seed(seed)
      conv = ref.layers.Conv1D(width=width, temporal_width=tw, dtype=dtype)
      with torch.no_grad():
        conv.w.copy_((torch.randn(conv.w.shape, generator=g) * 0.5).to(dtype))
        conv.b.copy_((torch.randn(conv.b.shape, generator=g) * 0.1).to(dtype))
      x = torch.randn((bsz, steps, width), generator=g).to(dtype)
      if segkind == "halves":
        seg = halves(steps, bsz)
      elif segkind == "ragged":
        seg = ragged_segments(bsz, steps, seed)
      elif segkind == "ragged_pad":
        seg = ragged_segments(bsz, steps, seed, pad=5)
      else:  # 1-D positions broadcast over the batch (layers.py:512-513)
        seg = halves(steps, 1)[0]
      with torch.no_grad():
        y, cache = conv(x.clone(), seg)  # clone: the reference mutates x (D2)
        out = dict(w=conv.w.data, b=conv.b.data, x=x, seg=seg, y=y,
                   cache=cache)
        # two decode steps with the returned cache (test_utils.py:92-107)
        for i in range(2):
          xs = torch.randn((bsz, 1, width), generator=g).to(dtype)
          ys, cache = conv(xs, seg[..., -1:] + 1 + i, cache)
          out[f"step{i}_x"], out[f"step{i}_y"] = xs, ys
          out[f"step{i}_cache"] = cache
      fixture_io.save(f"conv1d_{tag}_w{tw}_t{steps}_{segkind}", out)
  # bf16 activations with an fp32 cache (cache dtype wins, layers.py:483,:542)
  g = gen(77)
  conv = ref.layers.Conv1D(width=32, temporal_width=4, dtype=torch.bfloat16)
  with torch.no_grad():
    conv.w.copy_((torch.randn(conv.w.shape, generator=g) * 0.5).bfloat16())
    conv.b.copy_((torch.randn(conv.b.shape, generator=g) * 0.1).bfloat16())
    cache = torch.randn((2, 3, 32), generator=g)
    xs = torch.randn((2, 1, 32), generator=g).bfloat16()
    ys, new_cache = conv(xs, torch.full((2, 1), 7), cache)
  fixture_io.save("conv1d_bf16_w4_decode_f32cache", dict(
      w=conv.w.data, b=conv.b.data, cache_in=cache, step0_x=xs, step0_y=ys,
      step0_cache=new_cache))


def make_rglru(ref):
  for tag, dtype in DTYPES.items():
    for width, heads, steps, bsz, segkind in [
        (128, 1, 32, 1, "halves"), (256, 8, 64, 2, "halves"),
        (256, 8, 67, 2, "ragged_pad"), (128, 16, 128, 1, "halves"),
        (64, 2, 1, 3, "first"),
        # head widths 128 / 256: the shapes the fused tensor-core kernel takes
        (512, 2, 96, 2, "ragged_pad"), (256, 2, 70, 3, "halves"),
        (512, 2, 160, 1, "halves"),
    ]:
      if ONLY and not any(o in f"rglru_{tag}_e{width}_h{heads}_t{steps}_{segkind}" for o in ONLY):
        continue
      if width // heads >= 128 and heads > 1 and tag != "bf16":
        continue   # fused-path shapes: bf16 only (the fp32 copies would add 6 MB)
      seed = seed_of(tag, width, heads, steps, segkind)
      g = gen(seed)
      torch.manual_seed(seed)
      lru = ref.layers.RGLRU(width=width, num_heads=heads, dtype=dtype)
      with torch.no_grad():  # non-trivial gate biases (reference zero-inits)
        lru.input_gate.b.copy_(torch.randn(lru.input_gate.b.shape, generator=g).to(dtype))
        lru.a_gate.b.copy_(torch.randn(lru.a_gate.b.shape, generator=g).to(dtype))
      x = torch.randn((bsz, steps, width), generator=g).to(dtype)
      if segkind == "halves":
        seg = halves(steps, bsz)
      elif segkind == "first":
        seg = torch.zeros((bsz, steps), dtype=torch.int64)
      else:
        seg = ragged_segments(bsz, steps, seed, pad=3)
      with torch.no_grad():
        pre_x = lru.input_gate(x)
        pre_a = lru.a_gate(x)
        y, last_h = lru(x, seg)
        out = dict(a_param=lru.a_param.data, input_gate_w=lru.input_gate.w.data,
                   input_gate_b=lru.input_gate.b.data, a_gate_w=lru.a_gate.w.data,
                   a_gate_b=lru.a_gate.b.data, x=x, seg=seg, pre_x=pre_x,
                   pre_a=pre_a, y=y, last_h=last_h)
        cache = last_h
        for i in range(2):
          xs = torch.randn((bsz, 1, width), generator=g).to(dtype)
          out[f"step{i}_pre_x"] = lru.input_gate(xs)
          out[f"step{i}_pre_a"] = lru.a_gate(xs)
          ys, cache = lru(xs, seg[:, -1:] + 1 + i, cache)
          out[f"step{i}_x"], out[f"step{i}_y"] = xs, ys
          out[f"step{i}_last_h"] = cache
      fixture_io.save(f"rglru_{tag}_e{width}_h{heads}_t{steps}_{segkind}", out)


def make_recurrent_block(ref):
  # (fixture tag, dtype, width, heads, lru_width, steps); "bf16_h128" has head
  # width 128: the block whose RG-LRU (and gating product) runs on the fused
  # tensor-core kernel
  for tag, dtype, width, heads, lru_width, steps in [
      ("f32", torch.float32, 96, 4, 128, 48), ("bf16", torch.bfloat16, 96, 4, 128, 48),
      ("bf16_h128", torch.bfloat16, 64, 2, 256, 80)]:
    if ONLY and not any(o in f"recurrent_block_{tag}" for o in ONLY):
      continue
    seed = {"f32": 4242, "bf16": 4243}.get(tag, 4244)
    g = gen(seed)
    torch.manual_seed(seed)
    blk = ref.modules.RecurrentBlock(width=width, num_heads=heads, lru_width=lru_width,
                                     conv1d_temporal_width=4, dtype=dtype)
    with torch.no_grad():
      for p in (blk.rg_lru.input_gate.b, blk.rg_lru.a_gate.b, blk.conv_1d.b,
                blk.linear_x.bias, blk.linear_y.bias):
        p.copy_((torch.randn(p.shape, generator=g) * 0.3).to(dtype))
      blk.conv_1d.w.copy_((torch.randn(blk.conv_1d.w.shape, generator=g) * 0.4).to(dtype))
    bsz = 2
    x = torch.randn((bsz, steps, width), generator=g).to(dtype)
    seg = halves(steps, bsz)
    out = {"param." + k: v for k, v in blk.state_dict().items()}
    captured = {}
    def grab_in(m, args, kwargs):
      captured.setdefault("conv_in", kwargs["x"].clone())

    def grab_out(m, args, kwargs, res):
      captured.setdefault("rglru_out", res[0].clone())

    hooks = [
        blk.conv_1d.register_forward_pre_hook(grab_in, with_kwargs=True),
        blk.rg_lru.register_forward_hook(grab_out, with_kwargs=True),
    ]
    with torch.no_grad():
      y, cache = blk(x, seg)
      for h in hooks:
        h.remove()
      out.update(x=x, seg=seg, y=y, rg_lru_state=cache.rg_lru_state,
                 conv1d_state=cache.conv1d_state, **captured)
      for i in range(2):
        xs = torch.randn((bsz, 1, width), generator=g).to(dtype)
        ys, cache = blk(xs, seg[:, -1:] + 1 + i, cache)
        out[f"step{i}_x"], out[f"step{i}_y"] = xs, ys
        out[f"step{i}_rg_lru_state"] = cache.rg_lru_state
        out[f"step{i}_conv1d_state"] = cache.conv1d_state
    fixture_io.save(f"recurrent_block_{tag}", out)


def make_griffin_tiny():
  """BASELINE config 1 (tiny Griffin, fp32, CPU) at B=1 (quirk D4), T=128.

  Captures, for each recurrent block, the tensors that cross the hot-path
  boundary inside the real model: the Conv1D input, the RG-LRU output and both
  caches, plus the logits.  The GPU tests replay the captured Conv1D inputs
  through the CUDA path with the captured weights.
  """
  ref = ref_loader.load_reference_griffin()
  common = ref.common
  R, A = common.TemporalBlockType.RECURRENT, common.TemporalBlockType.ATTENTION
  cfg = common.GriffinConfig(
      vocab_size=1000, width=256, mlp_expanded_width=768, num_heads=8,
      lru_width=256, block_types=(R, R, A), embeddings_scale_by_sqrt_dim=True,
      attention_window_size=2048, logits_soft_cap=30.0)
  torch.manual_seed(0)
  import contextlib
  import io
  with contextlib.redirect_stdout(io.StringIO()):
    model = ref.griffin.Griffin(cfg, gradient_checkpointing=False,
                                dtype=torch.float32)
  model.eval()
  g = gen(0)
  with torch.no_grad():
    for blk in model.blocks:
      if hasattr(blk, "recurrent_block"):
        rb = blk.recurrent_block
        for p in (rb.rg_lru.input_gate.b, rb.rg_lru.a_gate.b, rb.conv_1d.b):
          p.copy_(torch.randn(p.shape, generator=g) * 0.5)
        rb.conv_1d.w.copy_(torch.randn(rb.conv_1d.w.shape, generator=g) * 0.4)
  steps = 128
  tokens = torch.randint(0, 1000, (1, steps), generator=g)
  seg = halves(steps, 1)
  out = dict(tokens=tokens, seg=seg)
  captured = {}
  hooks = []
  for i, blk in enumerate(model.blocks):
    if not hasattr(blk, "recurrent_block"):
      continue
    rb = blk.recurrent_block
    def grab_in(m, args, kwargs, i=i):
      captured[f"blk{i}_conv_in"] = kwargs["x"].clone()

    def grab_out(m, args, kwargs, res, i=i):
      captured[f"blk{i}_rglru_out"] = res[0].clone()

    hooks.append(rb.conv_1d.register_forward_pre_hook(grab_in, with_kwargs=True))
    hooks.append(rb.rg_lru.register_forward_hook(grab_out, with_kwargs=True))
    for k, v in rb.state_dict().items():
      if k.startswith(("conv_1d.", "rg_lru.")):
        out[f"blk{i}_param.{k}"] = v
  with torch.no_grad():
    logits, cache = model(tokens, seg, return_logits=True, return_cache=True)
  for h in hooks:
    h.remove()
  out.update(captured)
  out["logits"] = logits
  for key, c in cache.items():
    if hasattr(c, "rg_lru_state"):
      idx = key.split(".")[-1]
      out[f"blk{idx}_rg_lru_state"] = c.rg_lru_state
      out[f"blk{idx}_conv1d_state"] = c.conv1d_state
  fixture_io.save("griffin_tiny_f32_t128", out)


def make_grads(ref):
  """Gradients torch autograd derives from the reference's rnn_scan / RGLRU
  (the training path, training/train.py): pins the backward scan kernel
  (SURVEY.md section 8(f) row F4)."""
  for tag, dtype in DTYPES.items():
    for name, (bsz, steps, width, with_h0) in {
        "t37_h0": (2, 37, 48, True), "t37_noh0": (2, 37, 48, False),
        "t1_h0": (3, 1, 40, True), "t1_noh0": (3, 1, 40, False),
        "t300_h0": (2, 300, 72, True),
    }.items():
      g = gen(seed_of("grad", tag, name))
      x = torch.randn((bsz, steps, width), generator=g).to(dtype).requires_grad_()
      a = (0.5 + 0.4999 * torch.rand((bsz, steps, width), generator=g)).to(dtype).requires_grad_()
      reset = torch.rand((bsz, steps), generator=g) < 0.1
      reset[0, 0] = True
      h0 = torch.randn((bsz, width), generator=g).requires_grad_() if with_h0 else None
      gy = torch.randn((bsz, steps, width), generator=g).to(dtype)
      gh = torch.randn((bsz, width), generator=g)
      y, h_last = ref.layers.rnn_scan(x, a, reset, h0)
      torch.autograd.backward([y, h_last], [gy, gh])
      fixture_io.save(f"grad_rnn_scan_{tag}_{name}", dict(
          x=x, a=a, reset=reset, h0=h0, gy=gy, gh=gh, y=y, h_last=h_last,
          dx=x.grad, da=a.grad, dh0=None if h0 is None else h0.grad))
    for width, heads, steps, bsz, segkind, with_cache in [
        (128, 4, 48, 2, "halves", False), (256, 2, 70, 2, "ragged_pad", True)]:
      seed = seed_of("grad", tag, width, heads, steps, segkind)
      g = gen(seed)
      torch.manual_seed(seed)
      lru = ref.layers.RGLRU(width=width, num_heads=heads, dtype=dtype)
      with torch.no_grad():
        lru.input_gate.b.copy_(torch.randn(lru.input_gate.b.shape, generator=g).to(dtype))
        lru.a_gate.b.copy_(torch.randn(lru.a_gate.b.shape, generator=g).to(dtype))
      x = torch.randn((bsz, steps, width), generator=g).to(dtype).requires_grad_()
      seg = halves(steps, bsz) if segkind == "halves" else ragged_segments(bsz, steps, seed, pad=3)
      cache = torch.randn((bsz, width), generator=g).requires_grad_() if with_cache else None
      gy = torch.randn((bsz, steps, width), generator=g).to(dtype)
      gh = torch.randn((bsz, width), generator=g)
      y, last_h = lru(x, seg, cache)
      torch.autograd.backward([y, last_h], [gy, gh])
      out = dict(x=x, seg=seg, cache=cache, gy=gy, gh=gh, y=y, last_h=last_h, dx=x.grad,
                 dcache=None if cache is None else cache.grad)
      for k, v in lru.named_parameters():
        out["param." + k] = v.data
        out["grad." + k] = v.grad
      fixture_io.save(f"grad_rglru_{tag}_e{width}_h{heads}_t{steps}_{segkind}", out)


def make_grads_block(ref):
  """Conv1D and whole-RecurrentBlock gradients (autograd through the reference)."""
  for tag, dtype in DTYPES.items():
    for steps, width, bsz, segkind in [(33, 64, 3, "ragged"), (5, 32, 2, "halves"), (150, 64, 2, "ragged_pad")]:
      seed = seed_of("gradconv", tag, steps, segkind)
      g = gen(seed)
      torch.manual_seed(seed)
      conv = ref.layers.Conv1D(width=width, temporal_width=4, dtype=dtype)
      with torch.no_grad():
        conv.w.copy_((torch.randn(conv.w.shape, generator=g) * 0.5).to(dtype))
        conv.b.copy_((torch.randn(conv.b.shape, generator=g) * 0.1).to(dtype))
      x = torch.randn((bsz, steps, width), generator=g).to(dtype).requires_grad_()
      seg = {"halves": lambda: halves(steps, bsz), "ragged": lambda: ragged_segments(bsz, steps, seed),
             "ragged_pad": lambda: ragged_segments(bsz, steps, seed, pad=5)}[segkind]()
      gy = torch.randn((bsz, steps, width), generator=g).to(dtype)
      y, _ = conv(x * 1.0, seg)          # a non-leaf: the reference masks its input in place (D2)
      y.backward(gy)
      fixture_io.save(f"grad_conv1d_{tag}_t{steps}_{segkind}", dict(
          w=conv.w.data, b=conv.b.data, x=x, seg=seg, gy=gy, y=y, dx=x.grad, dw=conv.w.grad,
          db=conv.b.grad))
    seed = seed_of("gradblock", tag)
    g = gen(seed)
    torch.manual_seed(seed)
    blk = ref.modules.RecurrentBlock(width=64, num_heads=2, lru_width=256,
                                     conv1d_temporal_width=4, dtype=dtype)
    with torch.no_grad():
      for p in (blk.rg_lru.input_gate.b, blk.rg_lru.a_gate.b, blk.conv_1d.b,
                blk.linear_x.bias, blk.linear_y.bias):
        p.copy_((torch.randn(p.shape, generator=g) * 0.3).to(dtype))
      blk.conv_1d.w.copy_((torch.randn(blk.conv_1d.w.shape, generator=g) * 0.4).to(dtype))
    bsz, steps = 2, 80
    x = torch.randn((bsz, steps, 64), generator=g).to(dtype).requires_grad_()
    seg = halves(steps, bsz)
    gy = torch.randn((bsz, steps, 64), generator=g).to(dtype)
    y, _ = blk(x, seg)
    y.backward(gy)
    out = dict(x=x, seg=seg, gy=gy, y=y, dx=x.grad)
    for k, v in blk.named_parameters():
      out["param." + k] = v.data
      out["grad." + k] = v.grad
    fixture_io.save(f"grad_recurrent_block_{tag}", out)


def main():
  ref = ref_loader.load_reference()
  if ONLY == ["gradblock"]:
    make_grads_block(ref)
  elif ONLY == ["grad"]:
    make_grads(ref)
  elif ONLY:
    make_rglru(ref)
    make_recurrent_block(ref)
  else:
    make_rnn_scan(ref)
    make_conv1d(ref)
    make_rglru(ref)
    make_recurrent_block(ref)
    make_griffin_tiny()
    make_grads(ref)
    make_grads_block(ref)
  total = 0
  for f in sorted(os.listdir(fixture_io.GOLDEN_DIR)):
    if f.endswith(".npz"):
      total += os.path.getsize(os.path.join(fixture_io.GOLDEN_DIR, f))
  print(f"wrote {len(fixture_io.cases(''))} fixtures, {total/1e6:.2f} MB, "
        f"torch {torch.__version__}")


if __name__ == "__main__":
  main()
