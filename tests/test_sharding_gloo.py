"""World-size-2 gloo tests (CPU) of the batch-sharding host logic.

The per-shard computation is injected (the oracle torch port on CPU): the test
covers row partitioning, ragged shards, state gathering and that sharded ==
unsharded bit for bit (rows are independent), not the CUDA kernels.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cadence_gemma_b200.sharding import (BatchShardedHotPath, all_gather_rows,
                                         shard_bounds)


def test_shard_bounds_cover_batch():
  for batch in (0, 1, 2, 5, 8, 13, 256):
    for world in (1, 2, 3, 4, 8):
      rows = [shard_bounds(batch, world, r) for r in range(world)]
      assert rows[0][0] == 0 and rows[-1][1] == batch
      for (a, b), (c, d) in zip(rows, rows[1:]):
        assert b == c
      sizes = [b - a for a, b in rows]
      assert max(sizes) - min(sizes) <= 1


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


def _worker(rank, world, port, batch, ok):
  from oracle import torch_port
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    torch.set_num_threads(1)
    g = torch.Generator().manual_seed(3)
    width, heads, steps = 64, 2, 24
    lp = torch_port.init_rglru_params(width, heads, g)
    cw, cb = torch_port.init_conv_params(width, 4, g, w_scale=1.0)
    x = torch.randn((batch, steps, width), generator=g)
    seg = torch.arange(steps)[None].repeat(batch, 1)
    seg[:, 10:] -= 10

    def step(xs, segs, conv_cache, lru_cache):
      xc, conv_state = torch_port.conv1d_forward(cw, cb, xs, segs, conv_cache)
      y, last_h = torch_port.rglru_forward(lp, xc, segs, lru_cache)
      return y, last_h, conv_state

    runner = BatchShardedHotPath(step_fn=step)
    y_loc, h_all, c_all = runner.forward(x, seg)
    y_ref, h_ref, c_ref = step(x, seg, None, None)
    lo, hi = shard_bounds(batch, world, rank)
    assert torch.equal(y_loc, y_ref[lo:hi])
    assert h_all.shape == h_ref.shape and torch.equal(h_all, h_ref)
    assert torch.equal(c_all, c_ref)
    # decode continuation with the gathered caches, sharded again
    xs = torch.randn((batch, 1, width), generator=g)
    y1, h1, c1 = runner.forward(xs, seg[:, -1:] + 1, c_all, h_all)
    y1_ref, h1_ref, c1_ref = step(xs, seg[:, -1:] + 1, c_ref, h_ref)
    assert torch.equal(y1, y1_ref[lo:hi]) and torch.equal(h1, h1_ref) and torch.equal(c1, c1_ref)
    # ragged gather helper on its own
    t = torch.arange(batch * 3, dtype=torch.float32).view(batch, 3)
    assert torch.equal(all_gather_rows(t[lo:hi].clone(), batch), t)
    ok[rank] = 1
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [4, 5, 1])   # 1: rank 1 owns an empty shard
def test_sharded_equals_unsharded_world2(batch):
  world = 2
  ok = mp.get_context("spawn").Array("i", [0] * world)
  mp.spawn(_worker, args=(world, _free_port(), batch, ok), nprocs=world, join=True)
  assert list(ok) == [1] * world
