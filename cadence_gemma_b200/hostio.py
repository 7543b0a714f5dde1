"""Host-buffer entry point of the recurrent hot path.

``HostPrefill`` runs ``Conv1D.forward -> RGLRU.forward`` for activations that
live in pinned HOST memory and returns the results to pinned host memory.  The
batch rows of the path are independent recurrences (reference
``recurrentgemma/torch/layers.py:195-197`` works row-wise, the convolution is
depth-wise, ``:530-533``), so the batch is cut into row chunks and the three
legs -- host->device copy, kernels, device->host copy -- run on three CUDA
streams: chunk c+1 is uploaded and chunk c-1 is downloaded while chunk c is
computed.  PCIe is full duplex, so a step costs about max(H2D, D2H) instead of
H2D + kernels + D2H.

There is no CPU fallback: the arithmetic is the same ``Conv1D`` / ``RGLRU``
shims (sm_100a kernels through the C ABI).
"""
from __future__ import annotations

import torch

from cadence_gemma_b200 import pipeline


def bind_host_thread_to_device(device_index: int):
  """Binds the calling thread to the CPU cores next to GPU ``device_index`` (NVML's
  ideal CPU affinity), so that pinned host buffers allocated afterwards are placed
  on that GPU's memory node and the copy engines do not pull them across the
  socket interconnect -- with one process per GPU on a multi-socket host the
  host legs of ``HostPrefill`` otherwise share one socket's memory and one
  inter-socket link.  Returns the previous affinity set (restore it with
  ``os.sched_setaffinity(0, previous)``) or None if NVML is unavailable."""
  import os
  try:
    import pynvml
    pynvml.nvmlInit()
    bus = torch.cuda.get_device_properties(device_index).pci_bus_id
    handle = None
    for i in range(pynvml.nvmlDeviceGetCount()):
      h = pynvml.nvmlDeviceGetHandleByIndex(i)
      if pynvml.nvmlDeviceGetPciInfo(h).bus == bus:     # CUDA_VISIBLE_DEVICES remapping
        handle = h
        break
    if handle is None:
      handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
    previous = os.sched_getaffinity(0)
    pynvml.nvmlDeviceSetCpuAffinity(handle)
    return previous
  except Exception:   # no NVML / not permitted: placement stays with the OS default
    return None


class HostPrefill:
  """Pipelined prefill of one recurrent block's hot path from host buffers.

  Args:
    conv, lru: ``cadence_gemma_b200.Conv1D`` / ``RGLRU`` on a CUDA device.
    batch, steps: shape of the host activations ``[batch, steps, width]``.
    chunks: number of row chunks (must divide ``batch``).
    graph: capture the whole three-stream pipeline (copies included) into a
      CUDA graph on first use and replay it afterwards -- with many small
      chunks the ~40 asynchronous calls of a step cost more host time than the
      transfers take.  One graph per set of host buffers (re-captured when a
      different buffer is passed).
  """

  def __init__(self, conv, lru, batch: int, steps: int, chunks: int = 4, graph: bool = False):
    assert batch % chunks == 0, (batch, chunks)
    self.conv, self.lru = conv, lru
    self.batch, self.steps, self.chunks = batch, steps, chunks
    self.rows = batch // chunks
    dev = conv.w.device
    dtype = conv.w.dtype
    width = conv.width
    self.device = dev
    self.s_in = torch.cuda.Stream(dev)
    self.s_run = torch.cuda.Stream(dev)
    self.s_out = torch.cuda.Stream(dev)
    # Two device buffers per tensor are enough: a chunk's buffers are reused two
    # chunks later, after the consumer of their previous content has finished.
    # Everything is allocated here, once: no allocator traffic (and none of its
    # cross-stream bookkeeping) on the hot path.
    tw = conv.temporal_width if hasattr(conv, "temporal_width") else conv.w.shape[0]
    mk = lambda shape, dt: [torch.empty(shape, dtype=dt, device=dev) for _ in range(2)]
    self.x_dev = mk((self.rows, steps, width), dtype)
    self.seg_dev = mk((self.rows, steps), torch.int32)
    self.xc_dev = mk((self.rows, steps, width), dtype)
    self.y_dev = mk((self.rows, steps, width), dtype)
    self.h_dev = mk((self.rows, width), torch.float32)
    self.c_dev = mk((self.rows, tw - 1, width), dtype)
    self.ev_in = [torch.cuda.Event() for _ in range(2)]       # one per device-buffer slot
    self.ev_run = [torch.cuda.Event() for _ in range(2)]
    self.ev_out = [torch.cuda.Event() for _ in range(2)]
    self._g = 0                                               # chunks enqueued so far
    self.use_graph = graph
    self._graphs = {}
    self._streamed = False          # the previous call was a submit(): its chunks may still be in flight

  @torch.no_grad()
  def __call__(self, x_host, seg_host, y_host, h_host=None, cache_host=None):
    """``x_host [B,T,E]``, ``seg_host [B,T]`` int32 (pinned) -> ``y_host [B,T,E]``,
    ``h_host [B,E]`` fp32, ``cache_host [B,W-1,E]`` (pinned, optional).
    Returns once all work is ENQUEUED; the caller's current stream waits for
    the last download."""
    if not self.use_graph:
      return self._enqueue(x_host, seg_host, y_host, h_host, cache_host)
    key = tuple(None if t is None else t.data_ptr()
                for t in (x_host, seg_host, y_host, h_host, cache_host))
    g = self._graphs.get(key)
    if g is None:
      # warm up outside capture (workspaces, packed weights, lazy module loads)
      self._enqueue(x_host, seg_host, y_host, h_host, cache_host)
      torch.cuda.synchronize(self.device)
      g = torch.cuda.CUDAGraph()
      cap = torch.cuda.Stream(self.device)
      cap.wait_stream(torch.cuda.current_stream(self.device))
      with torch.cuda.graph(g, stream=cap):
        self._g = 0                                      # no dependency on events recorded outside the capture
        self._enqueue(x_host, seg_host, y_host, h_host, cache_host)
      torch.cuda.current_stream(self.device).wait_stream(cap)
      if len(self._graphs) > 8:
        self._graphs.clear()
      self._graphs[key] = g
    g.replay()
    return y_host

  @torch.no_grad()
  def submit(self, x_host, seg_host, y_host, h_host=None, cache_host=None):
    """Streaming form of ``__call__`` for back-to-back batches: enqueues the
    batch on the three pipeline streams WITHOUT joining them with the caller's
    stream, so the upload of batch n+1 runs under the kernels and the download
    of batch n (PCIe is full duplex: a batch then costs max(H2D, D2H), with no
    pipeline fill / drain per batch).  The device buffers are shared between
    consecutive batches; the reuse dependencies are carried by the per-slot
    events across calls.  Returns a CUDA event that fires when ``y_host`` /
    ``h_host`` / ``cache_host`` of THIS batch are complete
    (``event.synchronize()`` or ``stream.wait_event(event)``).  The host input
    buffers must already hold their data when this is called."""
    self._enqueue(x_host, seg_host, y_host, h_host, cache_host, join=False)
    done = torch.cuda.Event()
    done.record(self.s_out)
    return done

  def _enqueue(self, x_host, seg_host, y_host, h_host, cache_host, join=True):
    cur = torch.cuda.current_stream(self.device)
    if join:
      if self._streamed:                                 # drain submitted batches first
        cur.wait_stream(self.s_out)
        cur.wait_stream(self.s_run)
        self._streamed = False
      for s in (self.s_in, self.s_run, self.s_out):
        s.wait_stream(cur)
    # The g-th chunk ever enqueued works in device-buffer slot g & 1, across
    # calls: before a slot is overwritten, the kernels that read it (upload
    # side) and the download that read it (kernel side) must have finished --
    # the per-slot events carry that from one chunk, and one call, to the next.
    for c in range(self.chunks):
      r0, r1 = c * self.rows, (c + 1) * self.rows
      k = self._g & 1
      reuse = self._g >= 2
      self._g += 1
      with torch.cuda.stream(self.s_in):
        if reuse:
          self.s_in.wait_event(self.ev_run[k])           # kernels of the slot's last chunk have read x / seg
        self.x_dev[k].copy_(x_host[r0:r1], non_blocking=True)
        self.seg_dev[k].copy_(seg_host[r0:r1], non_blocking=True)
        self.ev_in[k].record(self.s_in)
      with torch.cuda.stream(self.s_run):
        self.s_run.wait_event(self.ev_in[k])
        if reuse:
          self.s_run.wait_event(self.ev_out[k])          # the slot's last results have been downloaded
        # ONE fused launch at RecurrentGemma shapes (pipeline.recurrent_hot_path);
        # xc_dev is only touched by the two-kernel route
        pipeline.recurrent_hot_path(self.conv, self.lru, self.x_dev[k], self.seg_dev[k],
                                    out=self.y_dev[k], last_h_out=self.h_dev[k],
                                    conv_out=self.xc_dev[k], conv_cache_out=self.c_dev[k])
        self.ev_run[k].record(self.s_run)
      with torch.cuda.stream(self.s_out):
        self.s_out.wait_event(self.ev_run[k])
        y_host[r0:r1].copy_(self.y_dev[k], non_blocking=True)
        if h_host is not None:
          h_host[r0:r1].copy_(self.h_dev[k], non_blocking=True)
        if cache_host is not None:
          cache_host[r0:r1].copy_(self.c_dev[k], non_blocking=True)
        self.ev_out[k].record(self.s_out)
    self._streamed = not join
    if join:
      cur.wait_stream(self.s_out)
      cur.wait_stream(self.s_run)
    return y_host
