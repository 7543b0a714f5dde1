"""Diagnoses the cost of the state all-gather next to the kernels (N >= 2)."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cadence_gemma_b200 as cg
import bench

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if len(sys.argv) > 1 and sys.argv[1] != "0":
  os.environ["NCCL_MAX_NCHANNELS"] = sys.argv[1]
dist.init_process_group("nccl", device_id=dev)
w = bench.WORKLOAD
host = bench.make_host_inputs(torch.bfloat16, seed=1 + rank)
conv = cg.Conv1D(w["width"], 4, device=dev, dtype=torch.bfloat16)
lru = cg.RGLRU(w["width"], w["heads"], device=dev, dtype=torch.bfloat16)
x, seg = host["x_lin"].to(dev), host["segment_pos"].to(dev)
gin = torch.zeros(w["batch"] * w["width"] * 4, device=dev)
gout = torch.zeros(gin.numel() * world, device=dev)
comm = torch.cuda.Stream(dev)
if os.environ.get("DIAG_SAMPLER"):
  smp = bench.ClockSampler(lr); smp.start()
EV = []
def step():
  if os.environ.get("DIAG_EVENTS"):
    for _ in range(4):
      e_ = torch.cuda.Event(enable_timing=True); e_.record(); EV.append(e_)
  xc, cs = conv(x, seg); y, h = lru(xc, seg)
  if os.environ.get("DIAG_COPIES"):
    global LAST
    LAST = (h, cs)
  return y, h
with torch.no_grad():
  for _ in range(5): step()
  with torch.cuda.stream(comm): dist.all_gather_into_tensor(gout, gin)
  torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
  for mode in ("same_stream", "side_stream"):
    gev, cpu = [], []
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(54):
      step()
      if (i + 1) % 18 == 0:
        t0 = time.perf_counter()
        if os.environ.get("DIAG_COPIES"):
          h_, cs_ = LAST
          gin[: h_.numel()].copy_(h_.view(-1)); gin[h_.numel():].copy_(cs_.float().view(-1))
        if mode == "same_stream":
          a = torch.cuda.Event(enable_timing=True); a.record()
          dist.all_gather_into_tensor(gout, gin)
          b = torch.cuda.Event(enable_timing=True); b.record()
        else:
          comm.wait_stream(torch.cuda.current_stream())
          with torch.cuda.stream(comm):
            a = torch.cuda.Event(enable_timing=True); a.record()
            dist.all_gather_into_tensor(gout, gin)
            b = torch.cuda.Event(enable_timing=True); b.record()
        cpu.append((time.perf_counter() - t0) * 1e6); gev.append((a, b))
    torch.cuda.current_stream().wait_stream(comm)
    e.record(); torch.cuda.synchronize()
    print(f"rank {rank} {mode}: {s.elapsed_time(e)/54*1e3:.1f} us/step; gather gpu us {[round(a.elapsed_time(b)*1e3) for a,b in gev]}; cpu us {[round(c) for c in cpu]}", flush=True)
    dist.barrier(); torch.cuda.synchronize()
dist.destroy_process_group()
