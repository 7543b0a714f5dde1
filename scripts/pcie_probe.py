"""PCIe ceiling of the box: pinned H2D, D2H and both at once (84 MB each way)."""
import torch
n = 8 * 2048 * 2560
xh = torch.empty(n, dtype=torch.bfloat16).pin_memory()
yh = torch.empty(n, dtype=torch.bfloat16).pin_memory()
xd = torch.empty(n, dtype=torch.bfloat16, device="cuda")
yd = torch.empty(n, dtype=torch.bfloat16, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, it=10):
  for _ in range(2):
    fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(it):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / it


def both():
  cur = torch.cuda.current_stream()
  s1.wait_stream(cur); s2.wait_stream(cur)
  with torch.cuda.stream(s1):
    xd.copy_(xh, non_blocking=True)
  with torch.cuda.stream(s2):
    yh.copy_(yd, non_blocking=True)
  cur.wait_stream(s1); cur.wait_stream(s2)


mb = n * 2 / 1e6
t = timed(lambda: xd.copy_(xh, non_blocking=True)); print(f"H2D {t:.3f} ms {mb / t:.1f} GB/s")
t = timed(lambda: yh.copy_(yd, non_blocking=True)); print(f"D2H {t:.3f} ms {mb / t:.1f} GB/s")
t = timed(both); print(f"both {t:.3f} ms {2 * mb / t:.1f} GB/s total")
