"""Benchmark of the recurrent hot path (Conv1D -> gate GEMMs -> RG-LRU scan).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1]): one RecurrentGemma-2B-shape recurrent
block's hot path, random init, bf16, batch 8, seq 2048 prefill, per GPU.  A
"step" is one pass of that batch through ``Conv1D.forward`` ->
``RGLRU.forward`` (block-diagonal gate GEMMs in cuBLAS + the fused gate/scan
kernel).  With N GPUs every rank runs its own batch of 8 (batch-sharded, weak
scaling, no collective inside the path; the small per-row states are
all-gathered on a side stream).

Prints ONE JSON line (see the task contract): metric = prefill tokens/s of the
hot path, with `roofline` for the dominant kernel (RG-LRU gate+scan, 4*s
algorithmic bytes per element), `cpu_baseline` (the oracle port of the
reference's eager torch path, timed on this box's host cores) and `e2e` (same
step driven from pinned HOST buffers, H2D + D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(batch=8, seq_len=2048, width=2560, heads=10, temporal_width=4)
GATHER_EVERY = int(os.environ.get("CG_BENCH_GATHER_EVERY", "18"))   # recurrent blocks per 2B prefill (common.py:90-101)
METRIC = "rglru_conv1d_prefill_tokens_per_sec"
# dram__bytes_read.sum + dram__bytes_write.sum of one fused-kernel launch at config 2
# (ncu --set full, profiles/r1_rglru_fused_kernel_ncu_summary.txt)
NCU_TRAFFIC_BYTES = 162999296   # 94.71 MB read + 68.29 MB written (profiles/r1_rglru_fused_kernel_ncu_summary.txt)
UNIT = "tokens/s"


def parse_args():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=50)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--impl", default="own", choices=["own", "reference"])
  ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
  ap.add_argument("--no-cpu-baseline", action="store_true")
  return ap.parse_args()


def measured_peak_gbs():
  path = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(path):
    with open(path) as f:
      return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
  return 6650.0, "fallback (B200_PROFILING.md)"


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
  """Samples SM clock and throttle reasons through NVML from a background
  thread while the timed region runs (nvidia-smi -lms is too slow to start for
  regions of a few milliseconds)."""

  REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40,
             "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

  def __init__(self, gpu_index):
    self.gpu_index = gpu_index
    self.samples, self.reasons = [], set()
    self.sm_max = None
    self._stop = threading.Event()
    self._thread = None
    self._err = None

  def _run(self):
    try:
      import pynvml
      pynvml.nvmlInit()
      # CUDA_VISIBLE_DEVICES remapping: match by PCI bus id of the torch device
      bus = torch.cuda.get_device_properties(self.gpu_index).pci_bus_id
      handle = None
      for i in range(pynvml.nvmlDeviceGetCount()):
        h = pynvml.nvmlDeviceGetHandleByIndex(i)
        if pynvml.nvmlDeviceGetPciInfo(h).bus == bus:
          handle = h
          break
      if handle is None:
        handle = pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)
      self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM)
      while not self._stop.is_set():
        self.samples.append(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM))
        mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(handle)
        for name, bit in self.REASONS.items():
          if mask & bit:
            self.reasons.add(name)
        time.sleep(0.002)
    except Exception as exc:  # pylint: disable=broad-except
      self._err = repr(exc)

  def start(self):
    self._thread = threading.Thread(target=self._run, daemon=True)
    self._thread.start()
    time.sleep(0.05)   # let NVML initialise before the timed region starts

  def stop(self):
    self._stop.set()
    if self._thread is not None:
      self._thread.join(timeout=5)
    out = {"sm_mhz": statistics.median(self.samples) if self.samples else None,
           "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
           "samples": len(self.samples)}
    if self._err:
      out["error"] = self._err
    return out


# ----------------------------------------------------------------------------
def run_cpu_port(host, steps, warmup, threads):
  """The oracle port of the reference's eager torch path on the host cores.

  This is the ONLY place bench.py touches oracle/: as the reported CPU
  baseline / the `--impl reference` arm, never on the product path.
  """
  from oracle import torch_port
  torch.set_num_threads(threads)
  params = torch_port.RGLRUParams(host["a_param"], host["input_gate_w"],
                                  host["input_gate_b"], host["a_gate_w"], host["a_gate_b"])
  times, out = [], None
  with torch.no_grad():
    for i in range(warmup + steps):
      t0 = time.perf_counter()
      xc, conv_state = torch_port.conv1d_forward(host["conv_w"], host["conv_b"],
                                                 host["x_lin"], host["segment_pos"])
      y, last_h = torch_port.rglru_forward(params, xc, host["segment_pos"])
      dt = time.perf_counter() - t0
      if i >= warmup:
        times.append(dt)
      out = (y, last_h, conv_state)
  return times, out


def reference_arm(args, dtype):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return   # rank 0 alone runs the CPU reference arm
  threads = os.cpu_count() or 1
  host = make_host_inputs(dtype)
  tokens = WORKLOAD["batch"] * WORKLOAD["seq_len"]
  times, _ = run_cpu_port(host, args.steps, args.warmup, threads)
  total = sum(times)
  value = tokens * len(times) / total
  sample = (f"full config-2 batch (B={WORKLOAD['batch']}, T={WORKLOAD['seq_len']}, "
            f"E={WORKLOAD['width']}) per step, {len(times)} steps after {args.warmup} warm-ups")
  line = {
      "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
      "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
      "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
      "config": config_dict(args, "cpu"),
      "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                       "sample": sample},
      "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
              "d2h_bytes_per_step": 0},
      "gpu_launches": 0,
  }
  print(json.dumps(line), flush=True)


def config_dict(args, where):
  w = WORKLOAD
  return {"workload": ("BASELINE configs[1]: RecurrentGemma-2B-shape RG-LRU+Conv1D block, "
                       f"random init, {args.dtype}, batch {w['batch']}, seq {w['seq_len']} prefill"),
          "batch_per_gpu": w["batch"], "global_batch": w["batch"] * args.gpus,
          "seq_len": w["seq_len"], "lru_width": w["width"], "num_heads": w["heads"],
          "conv1d_temporal_width": w["temporal_width"],
          "parallelism": (f"batch-sharded x{args.gpus}, no collective in the path; NCCL all-gather "
                          f"of last_h + conv cache once per {GATHER_EVERY} block steps (one 2B prefill)"),
          "l2": "working set 420 MB per step > 126 MB L2 (no explicit flush)",
          "clock_ramp": "W warm-up steps plus 400 ms of untimed steps before the timed region",
          "device": where}


# ----------------------------------------------------------------------------
def own_arm(args, dtype):
  import torch.distributed as dist
  import cadence_gemma_b200 as cg
  from cadence_gemma_b200 import _abi

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
  k_events = {"conv1d": [], "gate_gemm": [], "rglru": []}
  step_no = [0]

  def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e

  def step(x, seg, record=False, gather=True):
    if record:   # same calls as the module API makes, with events between the kernels
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
    else:
      xc, conv_state = conv(x, seg)
      y, last_h = lru(xc, seg)
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

  with torch.no_grad():
    # W warm-up steps, then keep stepping until the GPU has been busy for
    # CG_BENCH_RAMP_MS: a step is ~0.15 ms, so W steps alone end before the SM
    # clocks have ramped up from idle on a fresh box (seen: 2x slower steps)
    for _ in range(max(args.warmup, 3)):
      out = step(x_dev, seg_dev)
    torch.cuda.synchronize()
    ramp_ms = float(os.environ.get("CG_BENCH_RAMP_MS", "400"))
    t_ramp = time.perf_counter()
    while (time.perf_counter() - t_ramp) * 1e3 < ramp_ms:
      # wall-clock bounded, so the ranks run DIFFERENT numbers of these steps: no
      # collective inside (a gather here once left two ranks in different collectives)
      for _ in range(20):
        out = step(x_dev, seg_dev, gather=False)
      torch.cuda.synchronize()
    if world > 1:
      # exercise the whole gather path once outside the timed region: NCCL sets
      # connections up lazily and CUDA loads the small copy/cast kernels lazily
      # (tens of milliseconds on first use)
      gather_in[: out[1].numel()].copy_(out[1].view(-1))
      gather_in[out[1].numel():].copy_(out[2].float().view(-1))
      comm_stream.wait_stream(torch.cuda.current_stream())
      with torch.cuda.stream(comm_stream):
        dist.all_gather_into_tensor(gather_out, gather_in)
    step_no[0] = 0
    sync_all()

    # ---------------- device-resident timed region: EXACTLY K steps -------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = _abi.launch_count
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    # per-kernel CUDA events bracket every CG_BENCH_EVENT_EVERY-th step of the timed
    # region (each bracket costs ~1-2 us of stream time, so not every step)
    ev_every = max(1, min(args.steps, int(os.environ.get("CG_BENCH_EVENT_EVERY", "5"))))
    for i in range(args.steps):
      out = step(x_dev, seg_dev, record=(i % ev_every == ev_every // 2))
    if world > 1:
      torch.cuda.current_stream().wait_stream(comm_stream)
    t_end.record()
    sync_all()
    launches = _abi.launch_count - launches0
    clocks = sampler.stop()
    ms_total = t_start.elapsed_time(t_end)
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
      for ev in e2e_done:
        torch.cuda.current_stream().wait_event(ev)
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

  # max over ranks (device time)
  t = torch.tensor([ms_total, e2e_ms, k_us["rglru"], k_us["conv1d"], k_us["gate_gemm"]],
                   dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  ms_total, e2e_ms, k2_mean_us, conv_us, gemm_us = t.tolist()

  if rank == 0:
    peak, peak_src = measured_peak_gbs()
    k2_bytes = 4 * esize * nelem          # read x, gemm_x, gemm_a; write y
    achieved = k2_bytes / (k2_mean_us * 1e-6) / 1e9
    value = world * tokens * args.steps / (ms_total * 1e-3)
    e2e_value = world * tokens * e2e_steps / (e2e_ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": config_dict(args, "cuda"),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "ms_per_step": e2e_ms / e2e_steps,
                "how": (f"HostPrefill: {e2e_chunks} row chunks pipelined over upload / kernel / download "
                        "streams" + ("; consecutive steps streamed with submit() (upload of step n+1 under "
                                     "the download of step n)" if e2e_stream else ""))},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm",
                     "kernel": ("cg::fused::rglru_fused_kernel (tcgen05 gate GEMMs + gate math + scan in one "
                                "launch; events bracket the cg_rglru_fused_fwd call incl. its small prologue "
                                "launch)") if fused else
                               ("cg::scan_kernel (RG-LRU gates + scan; events bracket the cg_rglru_fwd call "
                                "incl. its small prologue launch)"),
                     "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": k2_bytes,
                     "algorithmic_bytes_note": ("SURVEY 8(d) K2 figure: 4 x 2 B per element (x, pre_x, pre_a "
                                                "in, y out).  The fused kernel keeps pre_x / pre_a in TMEM, "
                                                "so it moves only about half of that (see `traffic`) and is "
                                                "bound by the SM-wide issue / MUFU throughput of the epilogue (gate math "
                                                "+ scan: 51 instructions, 5 of them MUFU, per element; ablations in "
                                                "DESIGN.md section 9), not HBM"
                                                if fused else "SURVEY 8(d) K2 figure: 4 x s bytes per element"),
                     "us_per_launch": k2_mean_us, "traffic": NCU_TRAFFIC_BYTES if fused else None},
        "kernels_us": ({"conv1d": conv_us, "rglru_fused_tcgen05": k2_mean_us} if fused else
                       {"conv1d": conv_us, "gate_gemm_cublas_fused": gemm_us, "rglru": k2_mean_us}),
        "roofline_conv1d": {"bound": "hbm", "achieved": 2 * esize * nelem / (conv_us * 1e-6) / 1e9,
                            "peak": peak, "unit": "GB/s",
                            "frac": 2 * esize * nelem / (conv_us * 1e-6) / 1e9 / peak},
        "roofline_conv1d_plus_rglru": {   # whole block incl. the gate GEMMs, canonical 6 x s bytes
            "bound": "hbm", "achieved": 6 * esize * nelem / ((conv_us + gemm_us + k2_mean_us) * 1e-6) / 1e9,
            "peak": peak, "unit": "GB/s",
            "frac": 6 * esize * nelem / ((conv_us + gemm_us + k2_mean_us) * 1e-6) / 1e9 / peak},
        "arith_mode": cg.get_arith_mode(),
        "fused_tcgen05_rglru": bool(fused),
    }
    if world == 1 and not args.no_cpu_baseline:
      if prev_affinity is not None:     # the CPU baseline uses every host core again
        for tid in os.listdir("/proc/self/task"):       # worker threads inherit the bound mask
          try:
            os.sched_setaffinity(int(tid), prev_affinity)
          except OSError:
            pass
      threads = os.cpu_count() or 1
      host0 = make_host_inputs(dtype, seed=1)
      times, cpu_out = run_cpu_port(host0, 3, 1, threads)
      best = min(times)
      line["cpu_baseline"] = {
          "value": tokens / best, "unit": UNIT, "cores": threads, "kind": "port",
          "sample": "full config-2 batch (16384 tokens), best of 3 after 1 warm-up",
          "ms_per_step": best * 1e3}
      y_gpu = out[0].float().cpu()
      y_cpu = cpu_out[0].float()
      line["parity_vs_cpu_port"] = {
          "normwise": ((y_gpu - y_cpu).abs().max() / y_cpu.abs().max()).item(),
          "bit_identical_frac": (out[0].cpu() == cpu_out[0]).float().mean().item()}
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
