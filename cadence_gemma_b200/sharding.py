"""Batch sharding of the recurrent hot path over the GPUs of one node.

Every (batch row, channel) is an independent recurrence and the convolution is
depth-wise (reference layers.py:195-197, :530-533), so the path shards by batch
with **no collective inside it**: rank g of G owns a contiguous block of batch
rows of the activations, segment positions and caches; parameters are
replicated.  The only communication is one all-gather of the small per-row
outputs -- the RG-LRU state ``[B,E]`` fp32 and the Conv1D cache ``[B,W-1,E]`` --
when the caller wants a merged cache.  The activations ``y`` stay sharded
(gathering them would cost more than the kernels).

One process per GPU (``torchrun``); ``torch.distributed`` with NCCL on GPUs
(gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


def shard_bounds(batch: int, world_size: int, rank: int) -> tuple[int, int]:
  """Row range [lo, hi) of `rank`: contiguous, sizes differ by at most one."""
  assert 0 <= rank < world_size and batch >= 0
  base, extra = divmod(batch, world_size)
  lo = rank * base + min(rank, extra)
  return lo, lo + base + (1 if rank < extra else 0)


def shard_rows(t: Optional[torch.Tensor], world_size: int, rank: int):
  """This rank's rows of a batch-major tensor (a view; None passes through)."""
  if t is None:
    return None
  lo, hi = shard_bounds(t.shape[0], world_size, rank)
  return t[lo:hi]


def all_gather_rows(local: torch.Tensor, batch: int, group=None) -> torch.Tensor:
  """All-gathers row shards produced by `shard_bounds` into a [batch, ...] tensor.

  Shards may differ by one row; they are padded to the largest shard so a
  single ``all_gather_into_tensor`` (one NCCL kernel) moves everything.
  """
  world = dist.get_world_size(group)
  rank = dist.get_rank(group)
  rows_max = -(-batch // world)
  tail = tuple(local.shape[1:])
  send = local
  if local.shape[0] != rows_max:
    send = local.new_zeros((rows_max,) + tail)
    send[: local.shape[0]] = local
  recv = local.new_empty((world * rows_max,) + tail)
  dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
  if batch == world * rows_max:
    return recv
  parts = []
  for r in range(world):
    lo, hi = shard_bounds(batch, world, r)
    parts.append(recv[r * rows_max: r * rows_max + (hi - lo)])
  del rank
  return torch.cat(parts, dim=0)


class BatchShardedHotPath:
  """Runs Conv1D -> RG-LRU on this rank's batch rows and gathers the states.

  `step_fn(x, segment_pos, conv_cache, lru_cache) -> (y, last_h, conv_state)`
  is the per-shard computation; by default the CUDA modules given to the
  constructor.  (The CPU tests inject a CPU function here: the product path
  itself has no CPU fallback.)
  """

  def __init__(self, conv=None, lru=None, group=None,
               step_fn: Optional[Callable] = None):
    self.conv, self.lru, self.group = conv, lru, group
    self._step_fn = step_fn or self._cuda_step

  def _cuda_step(self, x, segment_pos, conv_cache, lru_cache):
    from cadence_gemma_b200 import pipeline   # the hot-path entry point (one launch where that is faster)
    y, conv_state, last_h = pipeline.recurrent_hot_path(self.conv, self.lru, x, segment_pos,
                                                        conv_cache=conv_cache, lru_cache=lru_cache)
    return y, last_h, conv_state

  def forward(self, x_full, segment_pos_full, conv_cache_full=None,
              lru_cache_full=None, gather_states: bool = True):
    """Inputs are full-batch tensors (every rank holds or can address them);
    returns (y_local, last_h, conv_state) with the states gathered to full
    batch if `gather_states`, else local."""
    world = dist.get_world_size(self.group) if dist.is_initialized() else 1
    rank = dist.get_rank(self.group) if dist.is_initialized() else 0
    batch = x_full.shape[0]
    seg = segment_pos_full
    if seg.ndim == 2 and seg.shape[0] == batch:
      seg = shard_rows(seg, world, rank)
    y, last_h, conv_state = self._step_fn(
        shard_rows(x_full, world, rank), seg,
        shard_rows(conv_cache_full, world, rank),
        shard_rows(lru_cache_full, world, rank))
    if gather_states and world > 1:
      last_h = all_gather_rows(last_h, batch, self.group)
      conv_state = all_gather_rows(conv_state, batch, self.group)
    return y, last_h, conv_state
