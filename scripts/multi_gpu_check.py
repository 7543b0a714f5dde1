"""Multi-GPU check of the batch-sharded hot path over NCCL; run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 scripts/multi_gpu_check.py > gpurun_out/r2_multi_gpu_nN.json

BASELINE config 4 (B=16, T=8192, 7 document starts per row) and config 5 (decode step, B=256,
caches warmed by a T=16 prefill) go through ``cadence_gemma_b200.sharding.BatchShardedHotPath``:
every rank computes its rows, ONE all-gather returns the merged caches, which are fed back
(sharded again) into a decode step.  Rank 0 also runs the whole batch alone and checks that
sharded == unsharded bit for bit (rows are independent recurrences); times are CUDA events,
max over ranks.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cadence_gemma_b200 as cg  # noqa: E402
from cadence_gemma_b200.sharding import BatchShardedHotPath, shard_bounds  # noqa: E402

WIDTH, HEADS = 2560, 10


def seg_with_resets(bsz, steps):
  seg = torch.arange(steps, dtype=torch.int32)[None].repeat(bsz, 1)
  for b in range(bsz):
    g = torch.Generator().manual_seed(3000 + b)
    for c in sorted((torch.randperm(steps - 1, generator=g)[:7] + 1).tolist()):
      seg[b, c:] = torch.arange(steps - c, dtype=torch.int32)
  return seg


def main():
  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  torch.manual_seed(3)          # identical parameters on every rank
  conv = cg.Conv1D(WIDTH, 4, device=dev, dtype=torch.bfloat16)
  lru = cg.RGLRU(WIDTH, HEADS, device=dev, dtype=torch.bfloat16)
  with torch.no_grad():
    conv.w.normal_(0, 0.4); conv.b.normal_(0, 0.2)
    lru.input_gate.b.normal_(); lru.a_gate.b.normal_()
  runner = BatchShardedHotPath(conv, lru)
  out = {"n_gpus": world}

  def timed(fn, iters):
    for _ in range(3):
      fn()
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
      fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], device=dev, dtype=torch.float64)
    if world > 1:
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

  with torch.no_grad():
    # ---------------- config 4: long-context prefill with document starts, B = 16
    bsz, steps = 16, 8192
    g = torch.Generator().manual_seed(3)
    x = torch.randn((bsz, steps, WIDTH), generator=g).to(torch.bfloat16).to(dev)
    seg = seg_with_resets(bsz, steps).to(dev)
    y_loc, h_all, c_all = runner.forward(x, seg)
    lo, hi = shard_bounds(bsz, world, rank)
    ok = True
    if rank == 0:
      y_ref, c_ref, h_ref = cg.recurrent_hot_path(conv, lru, x, seg)
      ok = bool(torch.equal(y_loc, y_ref[lo:hi]) and torch.equal(h_all, h_ref) and torch.equal(c_all, c_ref))
    ms = timed(lambda: runner.forward(x, seg), 10)
    ms_nogather = timed(lambda: runner.forward(x, seg, gather_states=False), 10)
    out["config4_b16_t8192"] = {"sharded_equals_unsharded": ok, "ms_per_step_with_gather": ms,
                                "ms_per_step_without_gather": ms_nogather,
                                "tokens_per_s_with_gather": bsz * steps / (ms * 1e-3),
                                "tokens_per_s_without_gather": bsz * steps / (ms_nogather * 1e-3)}
    del x, y_loc
    # ---------------- config 5: decode step, B = 256, caches from a T = 16 prefill
    bsz = 256
    x0 = torch.randn((bsz, 16, WIDTH), generator=g).to(torch.bfloat16).to(dev)
    seg0 = torch.arange(16, dtype=torch.int32)[None].repeat(bsz, 1).to(dev)
    _, h_all, c_all = runner.forward(x0, seg0)
    xs = torch.randn((bsz, 1, WIDTH), generator=g).to(torch.bfloat16).to(dev)
    pos = torch.full((bsz, 1), 16, dtype=torch.int32, device=dev)
    y1, h1, c1 = runner.forward(xs, pos, c_all, h_all)          # gathered caches fed back, sharded again
    ok = True
    if rank == 0:
      _, c_ref, h_ref = cg.recurrent_hot_path(conv, lru, x0, seg0)
      y1_ref, c1_ref, h1_ref = cg.recurrent_hot_path(conv, lru, xs, pos, conv_cache=c_ref, lru_cache=h_ref)
      lo, hi = shard_bounds(bsz, world, rank)
      ok = bool(torch.equal(y1, y1_ref[lo:hi]) and torch.equal(h1, h1_ref) and torch.equal(c1, c1_ref))
    us = 1e3 * timed(lambda: runner.forward(xs, pos, c_all, h_all), 50)
    us_ng = 1e3 * timed(lambda: runner.forward(xs, pos, c_all, h_all, gather_states=False), 50)
    out["config5_decode_b256"] = {"sharded_equals_unsharded": ok, "us_per_step_with_gather": us,
                                  "us_per_step_without_gather": us_ng,
                                  "tokens_per_s_without_gather": bsz / (us_ng * 1e-6),
                                  "note": "eager step of ONE block's hot path per rank (B/N rows), one-launch decode kernel"}
  if rank == 0:
    print(json.dumps(out))
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
