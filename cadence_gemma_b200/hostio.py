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


class HostPrefill:
  """Pipelined prefill of one recurrent block's hot path from host buffers.

  Args:
    conv, lru: ``cadence_gemma_b200.Conv1D`` / ``RGLRU`` on a CUDA device.
    batch, steps: shape of the host activations ``[batch, steps, width]``.
    chunks: number of row chunks (must divide ``batch``).
  """

  def __init__(self, conv, lru, batch: int, steps: int, chunks: int = 4):
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
    # two device staging buffers per leg are enough: a chunk's buffer is reused
    # two chunks later, after the consumer of its previous content has finished
    self.x_dev = [torch.empty((self.rows, steps, width), dtype=dtype, device=dev) for _ in range(2)]
    self.seg_dev = [torch.empty((self.rows, steps), dtype=torch.int32, device=dev) for _ in range(2)]
    self.ev_in = [torch.cuda.Event() for _ in range(chunks)]
    self.ev_run = [torch.cuda.Event() for _ in range(chunks)]
    self.ev_free = [torch.cuda.Event() for _ in range(chunks)]

  @torch.no_grad()
  def __call__(self, x_host, seg_host, y_host, h_host=None, cache_host=None):
    """``x_host [B,T,E]``, ``seg_host [B,T]`` int32 (pinned) -> ``y_host [B,T,E]``,
    ``h_host [B,E]`` fp32, ``cache_host [B,W-1,E]`` (pinned, optional).
    Returns once all work is ENQUEUED; the caller's current stream waits for
    the last download."""
    cur = torch.cuda.current_stream(self.device)
    for s in (self.s_in, self.s_run, self.s_out):
      s.wait_stream(cur)
    keep = []
    for c in range(self.chunks):
      r0, r1 = c * self.rows, (c + 1) * self.rows
      xb, sb = self.x_dev[c & 1], self.seg_dev[c & 1]
      with torch.cuda.stream(self.s_in):
        if c >= 2:
          self.s_in.wait_event(self.ev_free[c - 2])      # kernels of chunk c-2 have read the buffer
        xb.copy_(x_host[r0:r1], non_blocking=True)
        sb.copy_(seg_host[r0:r1], non_blocking=True)
        self.ev_in[c].record(self.s_in)
      with torch.cuda.stream(self.s_run):
        self.s_run.wait_event(self.ev_in[c])
        xc, conv_state = self.conv(xb, sb)
        y, last_h = self.lru(xc, sb)
        self.ev_free[c].record(self.s_run)
        self.ev_run[c].record(self.s_run)
      with torch.cuda.stream(self.s_out):
        self.s_out.wait_event(self.ev_run[c])
        y_host[r0:r1].copy_(y, non_blocking=True)
        if h_host is not None:
          h_host[r0:r1].copy_(last_h, non_blocking=True)
        if cache_host is not None:
          cache_host[r0:r1].copy_(conv_state, non_blocking=True)
      for t in (xc, conv_state, y, last_h):
        t.record_stream(self.s_out)
      keep.append((xc, y))
    cur.wait_stream(self.s_out)
    cur.wait_stream(self.s_run)
    return y_host
