"""Host link ceiling of the box: plain pinned-memory copies, no kernels.

Explains bench.py's `e2e` figure: a config-2 batch moves 84 MB up and 84 MB down per
step, so e2e tokens/s <= 16384 / (84 MB / duplex bandwidth per direction).  Run alone
(N=1) or under torchrun (all ranks copy at the same time, barrier before each leg) to
see how much of one GPU's host bandwidth is left when N GPUs of the box transfer
together.  Measurement tool only; prints one JSON line per rank-0 with the per-rank
minimum over ranks.

  python scripts/host_link_ceiling.py
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29571 scripts/host_link_ceiling.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  torch.cuda.set_device(local)
  if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
  nbytes = 8 * 2048 * 2560 * 2                      # one config-2 activation tensor
  up_h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
  dn_h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
  up_d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
  dn_d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
  s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
  reps = 20

  def leg(up, dn):
    for timed in (False, True):
      if world > 1:
        dist.barrier()
      torch.cuda.synchronize()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      cur = torch.cuda.current_stream()
      e0.record(cur)
      s_up.wait_event(e0)
      s_dn.wait_event(e0)
      for _ in range(reps):
        if up:
          with torch.cuda.stream(s_up):
            up_d.copy_(up_h, non_blocking=True)
        if dn:
          with torch.cuda.stream(s_dn):
            dn_h.copy_(dn_d, non_blocking=True)
      cur.wait_stream(s_up)
      cur.wait_stream(s_dn)
      e1.record(cur)
      torch.cuda.synchronize()
      ms = e0.elapsed_time(e1)
    gbs = nbytes * reps / (ms * 1e-3) / 1e9         # per direction
    t = torch.tensor([gbs], device="cuda")
    if world > 1:
      dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return float(t.item())

  out = {"n_gpus": world, "bytes_per_copy": nbytes,
         "h2d_only_GBps_per_gpu_min": leg(True, False),
         "d2h_only_GBps_per_gpu_min": leg(False, True),
         "duplex_GBps_per_direction_per_gpu_min": leg(True, True)}
  out["e2e_ceiling_tokens_per_s_all_gpus"] = (
      world * 8 * 2048 / (nbytes / (out["duplex_GBps_per_direction_per_gpu_min"] * 1e9)))
  if rank == 0:
    print(json.dumps(out))
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
