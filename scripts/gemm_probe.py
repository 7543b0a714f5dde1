"""Times the block-diagonal gate GEMM variants (cuBLAS via torch) at config 2."""
import torch
dev = "cuda:0"
B, T, E, H = 8, 2048, 2560, 10
bw = E // H
x = torch.randn(B * T, H, bw, device=dev, dtype=torch.bfloat16)
wx = torch.randn(H, bw, bw, device=dev, dtype=torch.bfloat16) * bw ** -0.5
wa = torch.randn(H, bw, bw, device=dev, dtype=torch.bfloat16) * bw ** -0.5
wcat = torch.cat([wx, wa], dim=2).contiguous()          # [H, bw, 2bw]
o1 = torch.empty(B * T, H, bw, device=dev, dtype=torch.bfloat16)
o2 = torch.empty_like(o1)
ocat = torch.empty(B * T, H, 2 * bw, device=dev, dtype=torch.bfloat16)

def timeit(fn, n=30):
  for _ in range(5): fn()
  torch.cuda.synchronize()
  s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  s.record()
  for _ in range(n): fn()
  e.record(); e.synchronize()
  return s.elapsed_time(e) * 1e3 / n

def two():
  torch.bmm(x.transpose(0, 1), wx, out=o1.transpose(0, 1))
  torch.bmm(x.transpose(0, 1), wa, out=o2.transpose(0, 1))
def one():
  torch.bmm(x.transpose(0, 1), wcat, out=ocat.transpose(0, 1))
def einsum_ref():
  return torch.einsum("nhi,hij->nhj", x, wx), torch.einsum("nhi,hij->nhj", x, wa)
# dense alternative: one [BT,E] x [E,2E] block-diagonal-as-dense GEMM would be 10x the flops: skip
print("two bmm (out=strided):", timeit(two), "us")
print("one bmm N=512 (out=strided):", timeit(one), "us")
print("two einsum (reference style):", timeit(einsum_ref), "us")
two(); one()
print("equal:", torch.equal(ocat[:, :, :bw], o1), torch.equal(ocat[:, :, bw:], o2))
# head-major layout variants (contiguous per head): is the strided output what hurts?
xh = x.transpose(0, 1).contiguous()
oh = torch.empty(H, B * T, 2 * bw, device=dev, dtype=torch.bfloat16)
print("one bmm head-major contiguous:", timeit(lambda: torch.bmm(xh, wcat, out=oh)), "us")
