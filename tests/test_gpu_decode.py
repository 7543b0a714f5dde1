"""Graphed decode step: ``GraphedRecurrentDecode`` replays one token
step of a stack of recurrent blocks as a CUDA graph, caches updated in place on the
device; it must give the same bits as stepping the modules eagerly."""
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.mark.parametrize("dtype,width,lru_width,heads,bsz", [
    (torch.bfloat16, 256, 512, 2, 5),       # fused shapes (head width 256)
    (torch.bfloat16, 128, 128, 4, 3),       # head width 32
    (torch.float32, 64, 128, 2, 2),
])
def test_graphed_decode_equals_eager_steps(dtype, width, lru_width, heads, bsz):
  from cadence_gemma_b200 import pipeline
  from cadence_gemma_b200.decode import GraphedRecurrentDecode
  from cadence_gemma_b200.modules import RecurrentBlock
  torch.manual_seed(width + heads)
  nblocks, prefill, nsteps = 3, 21, 6
  blocks = [RecurrentBlock(width=width, num_heads=heads, lru_width=lru_width, device=DEV, dtype=dtype).eval()
            for _ in range(nblocks)]
  with torch.no_grad():
    for blk in blocks:
      blk.rg_lru.input_gate.b.normal_(); blk.rg_lru.a_gate.b.normal_()
      blk.conv_1d.w.normal_(0, 0.4); blk.conv_1d.b.normal_(0, 0.2)
    x0 = torch.randn(bsz, prefill, width, device=DEV).to(dtype)
    seg = torch.arange(prefill, device=DEV, dtype=torch.int32)[None].repeat(bsz, 1)
    caches, x = [], x0
    for blk in blocks:                          # prefill warms the caches (h0 != 0)
      x, c = blk(x, seg)
      caches.append(c)
    xs = [torch.randn(bsz, 1, width, device=DEV).to(dtype) for _ in range(nsteps)]
    # eager reference: the modules, three small kernels per block (the graph's kernels)
    old = pipeline.set_fused_decode(False)
    try:
      want, cs = [], list(caches)
      for i, xt in enumerate(xs):
        pos = torch.full((bsz, 1), prefill + i, device=DEV, dtype=torch.int32)
        h = xt
        for j, blk in enumerate(blocks):
          h, cs[j] = blk(h, pos, cs[j])
        want.append(h.clone())
    finally:
      pipeline.set_fused_decode(old)
    dec = GraphedRecurrentDecode(blocks, caches, torch.full((bsz, 1), prefill, dtype=torch.int32))
    for i, xt in enumerate(xs):
      got = dec.step(xt)
      assert torch.equal(got, want[i]), (i, (got.float() - want[i].float()).abs().max().item())
    for c_got, c_want in zip(dec.caches, cs):
      assert torch.equal(c_got.rg_lru_state, c_want.rg_lru_state)
      assert torch.equal(c_got.conv1d_state, c_want.conv1d_state)
    assert int(dec.pos[0, 0]) == prefill + nsteps
  torch.cuda.synchronize()
