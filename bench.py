"""Benchmark of the recurrent hot path (Conv1D -> gate GEMMs -> RG-LRU scan).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload config2|config3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (default, BASELINE.json configs[1]): one RecurrentGemma-2B-shape
recurrent block's hot path, random init, bf16, batch 8, seq 2048 prefill, per
GPU.  A "step" is one pass of that batch through the hot-path entry point
``cadence_gemma_b200.recurrent_hot_path`` = ``Conv1D.forward -> RGLRU.forward``
(reference modules.py:638-649): the Conv1D kernel followed by the fused tcgen05
kernel (gate GEMMs + gate math + scan), or ONE launch with the convolution inside
that kernel where that is the faster route (pipeline.py).  With N GPUs every rank
runs its own batch of 8 (batch-sharded, weak scaling, no collective inside the
path; the small per-row states are all-gathered on a side stream once per 18
block steps = one 2B prefill).

``--workload config3`` (BASELINE.json configs[2]): the Cadence multimodal prefill
at MODEL level -- synthetic ViT features -> MLP projector -> RecurrentGemma-2B
stack (18 recurrent + 8 attention blocks), batch 32 sharded over the ranks, 256
visual + 512 text tokens (scripts/model_prefill.py).

Prints ONE JSON line (see the task contract): metric = prefill tokens/s of the
hot path, `roofline` for the dominant kernel on the bytes it really moves,
`roofline_step` on the canonical 12 B/element of SURVEY 8(d) (the number tracked
against the north star's 70 % target), `cpu_baseline` (the reference's own eager
torch path on this box's host cores), `parity` (bit-identical fraction against
that CPU run for the shipped and the exact arithmetic mode) and `e2e` (same step
driven from pinned HOST buffers, H2D + D2H inside the timed region, with the
box's host-link ceiling measured in the same process).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(batch=8, seq_len=2048, width=2560, heads=10, temporal_width=4)
GATHER_EVERY = int(os.environ.get("CG_BENCH_GATHER_EVERY", "18"))   # recurrent blocks per 2B prefill (common.py:90-101)
METRIC = "rglru_conv1d_prefill_tokens_per_sec"
UNIT = "tokens/s"
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full
# capture of the dominant kernel (not measurable inside an un-profiled run)
NCU_TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_traffic.json")


def parse_args():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=50)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--impl", default="own", choices=["own", "reference"])
  ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
  ap.add_argument("--workload", default="config2", choices=["config2", "config3"])
  ap.add_argument("--no-cpu-baseline", action="store_true")
  return ap.parse_args()


def measured_peaks():
  path = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(path):
    with open(path) as f:
      d = json.load(f)
    return float(d["hbm_gbs"]), float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
  return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


SAMPLE_REPLICAS = int(os.environ.get("CG_BENCH_SAMPLE_REPLICAS", "30"))   # the clock-sampled region is this many x K steps
LEAD_IN_STEPS = int(os.environ.get("CG_BENCH_LEAD_IN", "16"))   # untimed steps between the barrier and the start event


def ncu_traffic(kernel_key):
  try:
    with open(NCU_TRAFFIC_FILE) as f:
      d = json.load(f)
    e = d[kernel_key]
    return int(e["dram_bytes_per_launch"]), e["source"]
  except Exception:  # pylint: disable=broad-except
    return None, None


def make_host_inputs(dtype, seed=1):
  """Synthetic config-2 tensors + parameters on the host (deterministic)."""
  w = WORKLOAD
  g = torch.Generator().manual_seed(seed)
  bw = w["width"] // w["heads"]
  p = {
      "x_lin": torch.randn((w["batch"], w["seq_len"], w["width"]), generator=g).to(dtype),
      "segment_pos": torch.arange(w["seq_len"], dtype=torch.int32)[None].repeat(w["batch"], 1),
      "conv_w": (torch.randn((w["temporal_width"], w["width"]), generator=g) *
                 (0.01 / w["temporal_width"]) ** 0.5 * 10).to(dtype),
      "conv_b": (torch.randn((w["width"],), generator=g) * 0.1).to(dtype),
      "input_gate_w": (torch.randn((w["heads"], bw, bw), generator=g) * bw ** -0.5).to(dtype),
      "input_gate_b": torch.randn((w["heads"], bw), generator=g).to(dtype),
      "a_gate_w": (torch.randn((w["heads"], bw, bw), generator=g) * bw ** -0.5).to(dtype),
      "a_gate_b": torch.randn((w["heads"], bw), generator=g).to(dtype),
  }
  # a on the ring [0.9, 0.999] through the softplus parametrisation (layers.py:202-221)
  u = torch.empty(w["width"]).uniform_(0.9 ** 2 + 1e-8, 0.999 ** 2 + 1e-8, generator=g)
  p["a_param"] = torch.log(torch.exp(-0.5 * torch.log(u)) - 1.0).to(dtype)
  return p


# ----------------------------------------------------------------------------
# clocks / throttle sampling during the timed region
# ----------------------------------------------------------------------------
class ClockSampler:
  """SM clock and throttle reasons through NVML, sampled by the MAIN thread at points
  where the host has nothing to launch: between the batches of the clock-ramp loop and
  right after the last launch of a region has been enqueued while the GPU is still
  executing it (the host runs ahead of the GPU: `t_end.query()` is False for another
  millisecond or more).  A background thread polling NVML competes with the launch loop
  for the interpreter (and the driver) and was measured to slow the steps it samples."""

  REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40,
             "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

  def __init__(self, gpu_index):
    self.samples = []          # (tag, MHz, reason mask)
    self.sm_max = None
    self._err = None
    self._nvml = self._handle = None
    try:
      import pynvml
      pynvml.nvmlInit()
      # CUDA_VISIBLE_DEVICES remapping: match by PCI bus id of the torch device
      bus = torch.cuda.get_device_properties(gpu_index).pci_bus_id
      handle = None
      for i in range(pynvml.nvmlDeviceGetCount()):
        h = pynvml.nvmlDeviceGetHandleByIndex(i)
        if pynvml.nvmlDeviceGetPciInfo(h).bus == bus:
          handle = h
          break
      if handle is None:
        handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
      self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM)
      self._nvml, self._handle = pynvml, handle
    except Exception as exc:  # pylint: disable=broad-except
      self._err = repr(exc)

  def sample(self, tag):
    if self._handle is None:
      return
    try:
      mhz = self._nvml.nvmlDeviceGetClockInfo(self._handle, self._nvml.NVML_CLOCK_SM)
      mask = self._nvml.nvmlDeviceGetCurrentClocksEventReasons(self._handle)
      self.samples.append((tag, mhz, mask))
    except Exception as exc:  # pylint: disable=broad-except
      self._err = repr(exc)

  def sample_until(self, event, tag, limit=64):
    """Samples while `event` (recorded after the timed region's last launch) is pending."""
    n = 0
    while n < limit and not event.query():
      self.sample(tag)
      n += 1

  def summary(self):
    inside = [s for s in self.samples if s[0] == "timed"]
    ramp = [s for s in self.samples if s[0] == "ramp"]
    use = inside + ramp[-8:]   # (no "ramp" samples are taken any more; kept for tools that add them)
    reasons = set()
    for _, _, mask in use:
      for name, bit in self.REASONS.items():
        if mask & bit:
          reasons.add(name)
    out = {"sm_mhz": statistics.median(s[1] for s in use) if use else None,
           "sm_max_mhz": self.sm_max, "reasons": sorted(reasons),
           "samples": len(use), "samples_inside_timed_region": len(inside),
           "how": ("NVML polled from the main thread while the GPU executes a back-to-back REPLICA of the timed region "
                   "(the same steps, SAMPLE_REPLICAS x K of them; its step time is in ms_per_step_of_the_sampled_replica), so the timed region "
                   "itself carries no instrumentation")}
    if self._err:
      out["error"] = self._err
    return out


# ----------------------------------------------------------------------------
def cpu_reference_runner(host, threads):
  """Returns ``(kind, fn)``: ``fn()`` runs Conv1D.forward -> RGLRU.forward of the
  reference on the host cores and returns ``(y, last_h, conv_state)``.

  kind "reference": the reference's OWN classes, imported unmodified from
  baseline/_ref (or /root/reference); kind "port": oracle/torch_port.py, which
  issues the same ATen ops, when no copy of the reference is on the box.  This is
  the ONLY place bench.py touches oracle/: as the reported CPU baseline / the
  `--impl reference` arm, never on the product path.
  """
  torch.set_num_threads(threads)
  from oracle import ref_loader
  w = WORKLOAD
  if ref_loader.reference_available():
    ref = ref_loader.load_reference()
    dtype = host["x_lin"].dtype
    conv = ref.layers.Conv1D(width=w["width"], temporal_width=w["temporal_width"], dtype=dtype)
    lru = ref.layers.RGLRU(width=w["width"], num_heads=w["heads"], dtype=dtype)
    with torch.no_grad():
      conv.w.copy_(host["conv_w"]); conv.b.copy_(host["conv_b"])
      lru.a_param.copy_(host["a_param"])
      lru.input_gate.w.copy_(host["input_gate_w"]); lru.input_gate.b.copy_(host["input_gate_b"])
      lru.a_gate.w.copy_(host["a_gate_w"]); lru.a_gate.b.copy_(host["a_gate_b"])

    def run_ref():
      with torch.no_grad():
        # the reference masks its input in place (layers.py:524): give it a copy
        xc, conv_state = conv(host["x_lin"].clone(), host["segment_pos"])
        y, last_h = lru(xc, host["segment_pos"])
      return y, last_h, conv_state
    return "reference", run_ref

  from oracle import torch_port
  params = torch_port.RGLRUParams(host["a_param"], host["input_gate_w"],
                                  host["input_gate_b"], host["a_gate_w"], host["a_gate_b"])

  def run_port():
    with torch.no_grad():
      xc, conv_state = torch_port.conv1d_forward(host["conv_w"], host["conv_b"],
                                                 host["x_lin"], host["segment_pos"])
      y, last_h = torch_port.rglru_forward(params, xc, host["segment_pos"])
    return y, last_h, conv_state
  return "port", run_port


def time_cpu(fn, steps, warmup):
  times, out = [], None
  for i in range(warmup + steps):
    t0 = time.perf_counter()
    out = fn()
    dt = time.perf_counter() - t0
    if i >= warmup:
      times.append(dt)
  return times, out


def reference_arm(args, dtype):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return   # rank 0 alone runs the CPU reference arm
  if args.workload == "config3":
    from scripts import model_prefill
    print(json.dumps(model_prefill.reference_arm_line(args)), flush=True)
    return
  threads = os.cpu_count() or 1
  host = make_host_inputs(dtype)
  tokens = WORKLOAD["batch"] * WORKLOAD["seq_len"]
  kind, fn = cpu_reference_runner(host, threads)
  times, _ = time_cpu(fn, args.steps, args.warmup)
  total = sum(times)
  value = tokens * len(times) / total
  sample = (f"full config-2 batch (B={WORKLOAD['batch']}, T={WORKLOAD['seq_len']}, "
            f"E={WORKLOAD['width']}) per step, {len(times)} steps after {args.warmup} warm-ups")
  line = {
      "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
      "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
      "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
      "config": config_dict(args),   # the SAME config as the GPU arm
      "device": "cpu",
      "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                       "sample": sample,
                       "what": ("the reference's own Conv1D.forward + RGLRU.forward (unmodified, baseline/_ref)"
                                if kind == "reference" else "oracle/torch_port.py (same ATen ops as the reference)")},
      "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
              "d2h_bytes_per_step": 0},
      "gpu_launches": 0,
  }
  print(json.dumps(line), flush=True)


def config_dict(args):
  w = WORKLOAD
  return {"workload": ("BASELINE configs[1]: RecurrentGemma-2B-shape RG-LRU+Conv1D block, "
                       f"random init, {args.dtype}, batch {w['batch']}, seq {w['seq_len']} prefill"),
          "batch_per_gpu": w["batch"], "global_batch": w["batch"] * args.gpus,
          "seq_len": w["seq_len"], "lru_width": w["width"], "num_heads": w["heads"],
          "conv1d_temporal_width": w["temporal_width"],
          "parallelism": (f"batch-sharded x{args.gpus}, no collective in the path; NCCL all-gather "
                          f"of last_h + conv cache once per {GATHER_EVERY} block steps (one 2B prefill)"),
          "l2": ("working set 168 MB per step (x_lin in, y out; one-launch route) / 336 MB (two-kernel route) "
                 "> 126 MB L2 (no explicit flush)"),
          "clock_ramp": ("W warm-up steps plus 400 ms of untimed steps before the timed region; "
                         f"{LEAD_IN_STEPS} untimed lead-in steps between the barrier and the start event")}


def host_link_ceiling(dev, nbytes_up, nbytes_dn, reps=10):
  """Plain pinned-memory copies of one step's bytes, up and down at the same time,
  no kernels: the box's host-link ceiling for `e2e` (ms per step)."""
  up_h = torch.empty(nbytes_up, dtype=torch.uint8).pin_memory()
  dn_h = torch.empty(nbytes_dn, dtype=torch.uint8).pin_memory()
  up_d = torch.empty(nbytes_up, dtype=torch.uint8, device=dev)
  dn_d = torch.empty(nbytes_dn, dtype=torch.uint8, device=dev)
  s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
  ms = None
  for timed in (False, True):
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream(dev)
    e0.record(cur)
    s_up.wait_event(e0)
    s_dn.wait_event(e0)
    for _ in range(reps):
      with torch.cuda.stream(s_up):
        up_d.copy_(up_h, non_blocking=True)
      with torch.cuda.stream(s_dn):
        dn_h.copy_(dn_d, non_blocking=True)
    cur.wait_stream(s_up)
    cur.wait_stream(s_dn)
    e1.record(cur)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
  return ms


# ----------------------------------------------------------------------------
def own_arm(args, dtype):
  import torch.distributed as dist
  import cadence_gemma_b200 as cg
  from cadence_gemma_b200 import _abi, pipeline

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  if not torch.cuda.is_available():
    raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
  torch.cuda.set_device(local_rank)
  dev = torch.device("cuda", local_rank)
  if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    # the only collective moves ~400 KB of per-row state: two channels suffice
    # and keep NCCL's footprint on the SMs (shared with the kernels) small
    if os.environ.get("CG_BENCH_NCCL_CHANNELS", "2") != "0":
      os.environ.setdefault("NCCL_MAX_NCHANNELS", os.environ.get("CG_BENCH_NCCL_CHANNELS", "2"))
    import datetime
    # fail fast: a collective that does not complete is an error here, not a 10 minute wait
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
  _abi.load()

  if args.workload == "config3":
    from scripts import model_prefill
    line = model_prefill.run(args, dev, world, rank)
    if rank == 0:
      print(json.dumps(line), flush=True)
    if world > 1:
      dist.destroy_process_group()
    return

  w = WORKLOAD
  host = make_host_inputs(dtype, seed=1 + rank)   # every rank its own batch shard
  conv = cg.Conv1D(w["width"], w["temporal_width"], device=dev, dtype=dtype)
  lru = cg.RGLRU(w["width"], w["heads"], device=dev, dtype=dtype)
  conv.load_state_dict({"w": host["conv_w"], "b": host["conv_b"]})
  lru.load_state_dict({"a_param": host["a_param"], "input_gate.w": host["input_gate_w"],
                       "input_gate.b": host["input_gate_b"], "a_gate.w": host["a_gate_w"],
                       "a_gate.b": host["a_gate_b"]})
  x_dev = host["x_lin"].to(dev)
  seg_dev = host["segment_pos"].to(dev)
  tokens = w["batch"] * w["seq_len"]
  nelem = tokens * w["width"]
  esize = 2 if dtype == torch.bfloat16 else 4

  # small outputs gathered over NCCL on a side stream (never the activations)
  comm_stream = torch.cuda.Stream(dev) if world > 1 else None
  state_numel = w["batch"] * w["width"] * 4   # last_h (1) + conv cache (3 rows) as fp32 words
  gather_in = torch.empty(state_numel, dtype=torch.float32, device=dev) if world > 1 else None
  gather_out = (torch.empty(state_numel * world, dtype=torch.float32, device=dev)
                if world > 1 else None)

  fused = lru.uses_fused_kernel(x_dev)
  one_launch = pipeline.can_fuse_conv(conv, lru, x_dev)
  bufs = dict(out=torch.empty_like(x_dev), conv_out=torch.empty_like(x_dev),
              last_h_out=torch.empty((w["batch"], w["width"]), dtype=torch.float32, device=dev),
              conv_cache_out=torch.empty((w["batch"], w["temporal_width"] - 1, w["width"]), dtype=dtype, device=dev))
  k_events = {"conv1d": [], "gate_gemm": [], "rglru": []}
  step_no = [0]

  def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e

  def step(x, seg, record=False, gather=True):
    if record and not one_launch:
      # the same kernels the hot-path entry point launches, with events between them
      e0 = ev()
      xc, conv_state = conv(x, seg)
      e1 = ev()
      if fused:   # gate GEMMs + gates + scan: ONE tcgen05 kernel (no separate GEMM)
        e2 = e1
        y, last_h = lru(xc, seg)
      else:
        gates = lru.gate_gemm(xc)
        e2 = ev()
        y, last_h = _abi.rglru_fwd(xc, None, None, lru.input_gate.b, lru.a_gate.b, lru.a_param,
                                   seg, arith_mode=cg.get_arith_mode(), gemm_fused=gates,
                                   block_width=w["width"] // w["heads"])
      e3 = ev()
      k_events["conv1d"].append((e0, e1))
      if e2 is not e1:
        k_events["gate_gemm"].append((e1, e2))
      k_events["rglru"].append((e2, e3))
    elif record:
      e0 = ev()
      y, conv_state, last_h = cg.recurrent_hot_path(conv, lru, x, seg, **bufs)
      k_events["rglru"].append((e0, ev()))
    else:
      # caller-provided output buffers: no allocator traffic inside a step (a step that has to
      # cudaMalloc a fresh 84 MB block stalls the launch loop for ~1 ms: measured)
      y, conv_state, last_h = cg.recurrent_hot_path(conv, lru, x, seg, **bufs)
    step_no[0] += 1
    if gather and world > 1 and GATHER_EVERY > 0 and step_no[0] % GATHER_EVERY == 0:
      # merged cache for the host: one all-gather of the small per-row states
      # per model prefill (18 recurrent blocks), on a side stream
      gather_in[: last_h.numel()].copy_(last_h.view(-1))
      gather_in[last_h.numel():].copy_(conv_state.float().view(-1))
      comm_stream.wait_stream(torch.cuda.current_stream())
      with torch.cuda.stream(comm_stream):
        dist.all_gather_into_tensor(gather_out, gather_in)
    return y, last_h, conv_state

  def sync_all():
    if world > 1:
      torch.cuda.current_stream().wait_stream(comm_stream)
      torch.cuda.synchronize()
      dist.barrier()
    torch.cuda.synchronize()

  region_launches = [0]

  def timed_region(join_gather, sample_clocks=False, brackets=False, nsteps=None):
    """EXACTLY K steps between two events on this rank's stream.  join_gather: the
    side-stream all-gather is joined BEFORE the closing event (a consumer that needs
    the merged cache right away) or after it (the merged cache is only needed before
    the next decode step: SURVEY 8(e) asks for both figures)."""
    sync_all()
    # lead-in: the barrier above leaves the GPU idle for as long as the slowest rank takes to arrive
    # (0.1 - 1 ms at N > 1), and the first steps after an idle gap run slower (clocks, L2): a few UNTIMED
    # steps of the same kind go first, on the same stream, without any synchronisation -- the two events
    # below still bracket EXACTLY K steps (measured at N = 2: 134.3 us per step without, 128.8 with)
    for _ in range(LEAD_IN_STEPS):
      step(x_dev, seg_dev, record=False, gather=False)
    step_no[0] = 0
    region_launches[0] = _abi.launch_count                 # kernels launched between the two events only
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    o = None
    host_t = [time.perf_counter()]
    for i in range(args.steps if nsteps is None else nsteps):
      o = step(x_dev, seg_dev, record=brackets)
      if sample_clocks and i % 64 == 63:
        sampler.sample("timed")    # (the replica region only: the GPU has a queue of launched steps to work on meanwhile)
      host_t.append(time.perf_counter())
    if os.environ.get("CG_BENCH_DEBUG"):
      sys.stderr.write("host us per step: " + " ".join(f"{(b - a) * 1e6:.0f}" for a, b in zip(host_t, host_t[1:])) + "\n")
    if world > 1 and join_gather:
      torch.cuda.current_stream().wait_stream(comm_stream)
    t_end.record()
    region_launches[0] = _abi.launch_count - region_launches[0]
    if sample_clocks:
      sampler.sample_until(t_end, "timed")   # the GPU is still executing the region
    sync_all()
    return t_start.elapsed_time(t_end), o

  sampler = ClockSampler(local_rank)
  with torch.no_grad():
    # W warm-up steps, then keep stepping until the GPU has been busy for
    # CG_BENCH_RAMP_MS: a step is ~0.15 ms, so W steps alone end before the SM
    # clocks have ramped up from idle on a fresh box (seen: 2x slower steps)
    for _ in range(max(args.warmup, 3)):
      out = step(x_dev, seg_dev)
    # the bracketed variant of a step allocates its intermediates in a different pattern: let the
    # caching allocator see it BEFORE the timed region (a first-time cudaMalloc inside the region
    # cost 30 us per step: 177 us for the first region against 145 us for an identical second one)
    for _ in range(3):
      step(x_dev, seg_dev, record=True, gather=False)
    torch.cuda.synchronize()
    for v in k_events.values():
      v.clear()
    ramp_ms = float(os.environ.get("CG_BENCH_RAMP_MS", "400"))
    t_ramp = time.perf_counter()
    while (time.perf_counter() - t_ramp) * 1e3 < ramp_ms:
      # wall-clock bounded, so the ranks run DIFFERENT numbers of these steps: no
      # collective inside (a gather here once left two ranks in different collectives)
      for _ in range(20):
        out = step(x_dev, seg_dev, gather=False)
      torch.cuda.synchronize()   # (no NVML query in here: it idles the GPU for milliseconds and the
                                 # first region after such a gap was measured at 263 us per step, the next at 136)
    if world > 1:
      # exercise the whole gather path once outside the timed region: NCCL sets
      # connections up lazily and CUDA loads the small copy/cast kernels lazily
      # (tens of milliseconds on first use)
      gather_in[: out[1].numel()].copy_(out[1].view(-1))
      gather_in[out[1].numel():].copy_(out[2].float().view(-1))
      comm_stream.wait_stream(torch.cuda.current_stream())
      with torch.cuda.stream(comm_stream):
        dist.all_gather_into_tensor(gather_out, gather_in)

    # ---------------- device-resident timed region: EXACTLY K steps -------------
    ms_total, out = timed_region(join_gather=False)
    launches = region_launches[0]
    # clocks: the SAME K steps once more, right away, with NVML polled by the main thread while
    # the GPU executes them (after the last launch has been enqueued, so the polling cannot
    # delay a launch): the region that is timed carries no instrumentation at all, the region
    # that is sampled is its back-to-back replica (its own step time is reported beside it).
    # (SAMPLE_REPLICAS x K steps: an NVML query takes 1-2 ms, K steps only ~2.4 ms -- one sample; the replica is
    # long enough for a dozen)
    ms_sampled = timed_region(join_gather=False, sample_clocks=True, nsteps=args.steps * SAMPLE_REPLICAS)[0]
    # per-kernel times: the same K steps a third time with CUDA events between the kernels of
    # every step (the brackets cost stream time and a different call sequence on the host, so
    # they stay out of the region that is timed)
    timed_region(join_gather=False, brackets=True)
    clocks = sampler.summary()
    clocks["ms_per_step_of_the_sampled_replica"] = ms_sampled / (args.steps * SAMPLE_REPLICAS)
    clocks["steps_of_the_sampled_replica"] = args.steps * SAMPLE_REPLICAS
    ms_joined = timed_region(join_gather=True)[0] if world > 1 else ms_total
    k_us = {k: statistics.mean(a.elapsed_time(b) * 1e3 for a, b in v) if v else 0.0
            for k, v in k_events.items()}

    # ---------------- end to end: pinned host buffers in, results out ------------
    # the host buffers of this rank live on the memory node of its GPU
    from cadence_gemma_b200.hostio import bind_host_thread_to_device
    prev_affinity = (bind_host_thread_to_device(local_rank)
                     if os.environ.get("CG_BENCH_NUMA_BIND", "1") != "0" else None)
    x_pin = host["x_lin"].clone().pin_memory()
    seg_pin = host["segment_pos"].pin_memory()
    y_pin = torch.empty_like(x_pin).pin_memory()
    h_pin = torch.empty((w["batch"], w["width"]), dtype=torch.float32).pin_memory()
    c_pin = torch.empty((w["batch"], w["temporal_width"] - 1, w["width"]), dtype=dtype).pin_memory()
    h2d = x_pin.numel() * x_pin.element_size() + seg_pin.numel() * seg_pin.element_size()
    d2h = (y_pin.numel() * y_pin.element_size() + h_pin.numel() * 4 +
           c_pin.numel() * c_pin.element_size())

    # the public host-buffer entry point: row chunks pipelined over three streams
    # (upload c+1 | kernels c | download c-1), see cadence_gemma_b200/hostio.py
    from cadence_gemma_b200.hostio import HostPrefill
    e2e_chunks = int(os.environ.get("CG_BENCH_E2E_CHUNKS", "1"))
    e2e_graph = os.environ.get("CG_BENCH_E2E_GRAPH", "0") != "0"
    host_prefill = HostPrefill(conv, lru, w["batch"], w["seq_len"], chunks=e2e_chunks, graph=e2e_graph)

    # back-to-back batches are streamed (HostPrefill.submit): every step still
    # uploads its inputs and downloads its results inside the timed region, but
    # the upload of step n+1 overlaps the download of step n (PCIe full duplex).
    # Two sets of host result buffers alternate, as a double-buffered consumer would.
    e2e_stream = os.environ.get("CG_BENCH_E2E_STREAM", "1") != "0" and not e2e_graph
    outs = [(y_pin, h_pin, c_pin),
            (torch.empty_like(y_pin).pin_memory(), torch.empty_like(h_pin).pin_memory(),
             torch.empty_like(c_pin).pin_memory())]
    e2e_done = []

    def e2e_step():
      yo, ho, co = outs[step_no[0] & 1]
      if e2e_stream:
        e2e_done.append(host_prefill.submit(x_pin, seg_pin, yo, ho, co))
      else:
        host_prefill(x_pin, seg_pin, yo, ho, co)
      step_no[0] += 1

    def e2e_join():                                      # results of every submitted step are on the host
      for e in e2e_done:
        torch.cuda.current_stream().wait_event(e)
      e2e_done.clear()

    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(3):
      e2e_step()
    e2e_join()
    sync_all()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    for _ in range(e2e_steps):
      e2e_step()
    e2e_join()
    if world > 1:
      torch.cuda.current_stream().wait_stream(comm_stream)
    e_end.record()
    sync_all()
    e2e_ms = e_start.elapsed_time(e_end)
    # host-link ceiling of this box for the same bytes, all ranks copying at the same time
    sync_all()
    link_ms = host_link_ceiling(dev, h2d, d2h)

  # max over ranks (device time)
  t = torch.tensor([ms_total, e2e_ms, k_us["rglru"], k_us["conv1d"], k_us["gate_gemm"], ms_joined, link_ms],
                   dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  ms_total, e2e_ms, k2_mean_us, conv_us, gemm_us, ms_joined, link_ms = t.tolist()

  if rank == 0:
    peak, peak_tf, peak_src = measured_peaks()
    value = world * tokens * args.steps / (ms_total * 1e-3)
    e2e_value = world * tokens * e2e_steps / (e2e_ms * 1e-3)
    step_us = 1e3 * ms_total / args.steps
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": config_dict(args),
        "device": "cuda",
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "ms_per_step": e2e_ms / e2e_steps,
                "how": (f"HostPrefill: {e2e_chunks} row chunks pipelined over upload / kernel / download "
                        "streams" + ("; consecutive steps streamed with submit() (upload of step n+1 under "
                                     "the download of step n)" if e2e_stream else "")),
                "ceiling": {"value": world * tokens / (link_ms * 1e-3), "unit": UNIT,
                            "ms_per_step": link_ms,
                            "what": ("host-link ceiling of THIS box measured in this process: plain pinned copies of "
                                     "one step's H2D and D2H bytes at the same time on every rank, no kernels; e2e "
                                     "cannot exceed it and its scaling over N is the box's host links, not the path"),
                            "e2e_frac_of_ceiling": link_ms / (e2e_ms / e2e_steps)}},
        "gpu_launches": launches,
        "launches_per_step": launches / args.steps,
        "route": ("one launch: conv inside the fused tcgen05 kernel (cg_recurrent_prefill_fwd)" if one_launch else
                  "conv1d kernel + fused tcgen05 kernel (cg_conv1d_fwd, cg_rglru_fused_fwd)" if fused else
                  "conv1d kernel + cuBLAS gate GEMM + scan kernel"),
        "arith_mode": cg.get_arith_mode(),
        "fused_tcgen05_rglru": bool(fused),
    }
    if world > 1:
      line["with_gather_joined"] = {
          "value": world * tokens * args.steps / (ms_joined * 1e-3), "unit": UNIT,
          "ms_per_step": ms_joined / args.steps,
          "what": ("the same K steps with the side-stream all-gather of last_h + conv cache JOINED before the closing "
                   "event; `value` enqueues the same gather but joins it after the closing event (the merged cache is "
                   "only needed before the next decode step) -- SURVEY 8(e): both figures")}
    # ---- roofline of the dominant kernel, on the bytes it really moves
    if fused:
      wbytes = 2 * w["width"] * (w["width"] // w["heads"]) * esize     # both gates' block-diagonal weights, once
      io_elems = 2 if not one_launch else 2      # x in, y out (pre-activations stay in TMEM)
      k_bytes = io_elems * esize * nelem + wbytes
      flop = 2 * 2 * tokens * w["width"] * (w["width"] // w["heads"])  # two gate GEMMs
      traffic, traffic_src = ncu_traffic("fused_conv" if one_launch else "fused")
      line["roofline"] = {
          "bound": "hbm",
          "kernel": ("cg::fused::rglru_fused_kernel<CONV> (Conv1D + tcgen05 gate GEMMs + gate math + scan)"
                     if one_launch else
                     "cg::fused::rglru_fused_kernel (tcgen05 gate GEMMs + gate math + scan in one launch); events "
                     "bracket the cg_rglru_fused_fwd call incl. its ~2 us prologue launch"),
          "achieved": k_bytes / (k2_mean_us * 1e-6) / 1e9, "peak": peak, "unit": "GB/s",
          "frac": k_bytes / (k2_mean_us * 1e-6) / 1e9 / peak, "peak_source": peak_src,
          "algorithmic_bytes_per_launch": k_bytes,
          "algorithmic_bytes_note": ("what this kernel has to move: x in + y out = 2 x 2 B per element, plus the packed "
                                     "gate weights once (the gate pre-activations never leave TMEM)"),
          "us_per_launch": k2_mean_us, "traffic": traffic, "traffic_source": traffic_src,
          "limiter": ("NOT HBM: the epilogue's issue / FMA-pipe / XU rate (gate math with every bf16 rounding point of "
                      "the reference + scan; DESIGN.md section 9) and, close behind, the tensor side at N = 64"),
          "tensor": {"useful_flop_per_launch": flop, "achieved": flop / (k2_mean_us * 1e-6) / 1e12,
                     "peak": peak_tf, "unit": "TFLOP/s", "frac": flop / (k2_mean_us * 1e-6) / 1e12 / peak_tf}}
    else:
      k_bytes = 4 * esize * nelem
      line["roofline"] = {"bound": "hbm", "kernel": "cg::scan_kernel (RG-LRU gates + scan)",
                          "achieved": k_bytes / (k2_mean_us * 1e-6) / 1e9, "peak": peak, "unit": "GB/s",
                          "frac": k_bytes / (k2_mean_us * 1e-6) / 1e9 / peak, "peak_source": peak_src,
                          "algorithmic_bytes_per_launch": k_bytes, "us_per_launch": k2_mean_us, "traffic": None}
    line["kernels_us"] = ({"recurrent_prefill_fused_tcgen05": k2_mean_us} if one_launch else
                          {"conv1d": conv_us, "rglru_fused_tcgen05": k2_mean_us} if fused else
                          {"conv1d": conv_us, "gate_gemm_cublas_fused": gemm_us, "rglru": k2_mean_us})
    if conv_us > 0:
      line["roofline_conv1d"] = {"bound": "hbm", "achieved": 2 * esize * nelem / (conv_us * 1e-6) / 1e9,
                                 "peak": peak, "unit": "GB/s",
                                 "frac": 2 * esize * nelem / (conv_us * 1e-6) / 1e9 / peak}
    # ---- the number tracked against the north star: the whole step on the canonical bytes
    canon = 6 * esize * nelem
    line["roofline_step"] = {
        "bound": "hbm", "what": ("Conv1D + gate GEMMs + RG-LRU on SURVEY 8(d)'s canonical 6 x s = 12 B per element "
                                 "(read x_lin; write+read x_conv, pre_x, pre_a; write y) over ms_per_step"),
        "algorithmic_bytes_per_step": canon, "achieved": canon / (step_us * 1e-6) / 1e9, "peak": peak,
        "unit": "GB/s", "frac": canon / (step_us * 1e-6) / 1e9 / peak, "target_frac": 0.70,
        "us_per_step": step_us, "us_at_target": canon / (0.70 * peak * 1e9) * 1e6}
    if world == 1 and not args.no_cpu_baseline:
      if prev_affinity is not None:     # the CPU baseline uses every host core again
        for tid in os.listdir("/proc/self/task"):       # worker threads inherit the bound mask
          try:
            os.sched_setaffinity(int(tid), prev_affinity)
          except OSError:
            pass
      threads = os.cpu_count() or 1
      host0 = make_host_inputs(dtype, seed=1)
      kind, fn = cpu_reference_runner(host0, threads)
      times, cpu_out = time_cpu(fn, 3, 1)
      best = min(times)
      line["cpu_baseline"] = {
          "value": tokens / best, "unit": UNIT, "cores": threads, "kind": kind,
          "sample": "full config-2 batch (16384 tokens), best of 3 after 1 warm-up",
          "ms_per_step": best * 1e3}
      # ---- parity of the whole step against that CPU run, shipped mode and exact mode
      def parity_of(mode):
        old = cg.set_arith_mode(mode)
        try:
          with torch.no_grad():
            for _ in range(3):
              y, _, h = cg.recurrent_hot_path(conv, lru, x_dev, seg_dev, **bufs)
            torch.cuda.synchronize()
            e0 = ev()
            for _ in range(10):
              cg.recurrent_hot_path(conv, lru, x_dev, seg_dev, **bufs)
            e1 = ev()
            torch.cuda.synchronize()
        finally:
          cg.set_arith_mode(old)
        yg, yc = y.float().cpu(), cpu_out[0].float()
        return {"arith_mode": mode, "us_per_step": e0.elapsed_time(e1) * 100.0,
                "bit_identical_frac": (y.cpu() == cpu_out[0]).float().mean().item(),
                "normwise": ((yg - yc).abs().max() / yc.abs().max()).item(),
                "within_rtol1e-2_atol3e-2": bool(torch.allclose(yg, yc, rtol=1e-2, atol=3e-2)),
                "last_h_normwise": ((h.cpu() - cpu_out[1]).abs().max() / cpu_out[1].abs().max()).item()}
      line["parity"] = {"against": f"cpu_baseline ({kind}) on the same inputs, y of the whole step",
                        "shipped": parity_of(cg.get_arith_mode()),
                        "exact_reference": parity_of(_abi.ARITH_REFERENCE)}
    print(json.dumps(line), flush=True)
  if world > 1:
    dist.destroy_process_group()


def main():
  args = parse_args()
  dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
  if args.impl == "reference":
    reference_arm(args, dtype)
  else:
    own_arm(args, dtype)


if __name__ == "__main__":
  main()
