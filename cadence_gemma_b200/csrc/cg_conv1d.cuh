// Depth-wise causal temporal convolution for sm_100a.
// Replaces reference recurrentgemma/torch/layers.py:458-546 (Conv1D.forward):
// prefill with the fork's document mask (:592-633), decode step (:478-483),
// returned cache (:541-546, :650-662).
#pragma once

#include "cg_common.cuh"
#include "cg_scan.cuh"   // IoVec, load_io / store_io

#ifndef CG_CONV_ST
#define CG_CONV_ST 0
#endif

namespace cg {

struct ConvParams {
  const void* x;      // [B,T,E]
  const void* w;      // [W,E]
  const void* bias;   // [E]
  const void* seg;    // prefill only
  long long seg_bstride;
  int seg_is_i64;
  void* y;            // [B,T,E]
  void* cache_out;    // [B,W-1,E] or null
  const void* cache_in;   // decode only
  int cache_is_bf16;      // decode only
  int B, T, E, W;
  int mask_mode;
};

// ---------------------------------------------------------------------------
// W == 4 prefill.  Thread = 16-byte channel vector x LC consecutive steps with
// a 3-row register window; 8 lanes cover one 128 B row, lane groups / warps
// cover consecutive time tiles.  EMUL: accumulate in the tensor dtype with one
// rounding per eager op of the reference (:533-536) -- for bf16 these are
// packed HMUL2/HADD2, for fp32 non-contracted FMUL/FADD, both bit-exact with
// the reference.  !EMUL (bf16 only): fp32 accumulation, one final rounding.
// ---------------------------------------------------------------------------
// One tile = 128 threads: channel tile `ctile` (8 lanes x 16 B), 16 time slots of
// LC steps starting at slot block `tblock`, batch row b.
template <typename IO, bool EMUL, int LC>
__device__ __forceinline__ void conv1d_w4_tile(const ConvParams& p, int ctile, int tblock, int b) {
  constexpr int V = IoVec<IO>::V;
  constexpr bool BF = IoVec<IO>::kBf16;
  constexpr int EC = kCvl * V;
  const int cv = threadIdx.x & 7;
  const int tslot = tblock * (blockDim.x >> 3) + (threadIdx.x >> 3);
  const int ch0 = ctile * EC + cv * V;
  const int t0 = tslot * LC;
  if (ch0 >= p.E || t0 >= p.T) return;
  const IO* xb = reinterpret_cast<const IO*>(p.x) + (size_t)b * p.T * p.E + ch0;
  IO* yb = reinterpret_cast<IO*>(p.y) + (size_t)b * p.T * p.E + ch0;
  const long long seg0 = (long long)b * p.seg_bstride;

  // taps: wk[k] multiplies x[t - (3 - k)]  (w[W-1-shift], :530)
  uint4 wk[4], bv;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    wk[k] = *reinterpret_cast<const uint4*>(reinterpret_cast<const IO*>(p.w) + (size_t)k * p.E + ch0);
  bv = *reinterpret_cast<const uint4*>(reinterpret_cast<const IO*>(p.bias) + ch0);

  // All LC + 3 input rows of this thread's tile are requested before any
  // arithmetic (memory-level parallelism: the kernel is pure streaming).
  const uint4 zero = make_uint4(0, 0, 0, 0);
  uint4 xr[LC + 3];            // xr[r] = x[t0 - 3 + r]
#pragma unroll
  for (int r = 0; r < LC + 3; ++r) {
    const int t = t0 - 3 + r;
    xr[r] = (t >= 0 && t < p.T) ? ldg_stream(xb + (size_t)t * p.E) : zero;
  }
  // nb bit r = (segment_pos[t0 - 2 + r] != 0); positions before 0 are never
  // consulted for a tap that exists, so their value is irrelevant.
  unsigned nb = 0;
#pragma unroll
  for (int r = 0; r < LC + 2; ++r) {
    const int t = t0 - 2 + r;
    if (t >= 0 && t < p.T)
      nb |= (load_seg(p.seg, p.seg_is_i64 != 0, seg0 + t) != 0 ? 1u : 0u) << r;
  }
  const uint32_t k0[4] = {wk[0].x, wk[0].y, wk[0].z, wk[0].w}, k1[4] = {wk[1].x, wk[1].y, wk[1].z, wk[1].w};
  const uint32_t k2[4] = {wk[2].x, wk[2].y, wk[2].z, wk[2].w}, k3[4] = {wk[3].x, wk[3].y, wk[3].z, wk[3].w};
  const uint32_t bb[4] = {bv.x, bv.y, bv.z, bv.w};

#pragma unroll
  for (int j = 0; j < LC; ++j) {
    const int t = t0 + j;
    if (t >= p.T) break;
    // document mask per tap (shift s looks at segment_pos[t-s+1 .. ])
    const unsigned w3 = nb >> j;       // bit 0: seg[t-2], bit 1: seg[t-1], bit 2: seg[t]
    bool m1 = true, m2 = true, m3;
    if (p.mask_mode == 0) {
      m3 = (w3 & 1u) != 0;                        // fork: only seg[t-2], :629-632
    } else {
      m1 = (w3 & 4u) != 0;                        // upstream: seg[t-s+1..t] all != 0
      m2 = (w3 & 6u) == 6u;
      m3 = (w3 & 7u) == 7u;
    }
    const uint4 x0 = xr[j + 3];
    const uint4 a1 = m1 ? xr[j + 2] : zero, a2 = m2 ? xr[j + 1] : zero, a3 = m3 ? xr[j] : zero;
    const uint32_t s0[4] = {x0.x, x0.y, x0.z, x0.w}, s1[4] = {a1.x, a1.y, a1.z, a1.w};
    const uint32_t s2[4] = {a2.x, a2.y, a2.z, a2.w}, s3[4] = {a3.x, a3.y, a3.z, a3.w};
    uint32_t o[4];
    if constexpr (BF && EMUL) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t acc = bf2_mul(s0[i], k3[i]);                 // shift 0
        acc = bf2_add(acc, bf2_mul(s1[i], k2[i]));            // shift 1
        acc = bf2_add(acc, bf2_mul(s2[i], k1[i]));            // shift 2
        acc = bf2_add(acc, bf2_mul(s3[i], k0[i]));            // shift 3
        o[i] = bf2_add(acc, bb[i]);                           // + b, :536
      }
    } else if constexpr (BF) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float lo = bf_lo(s0[i]) * bf_lo(k3[i]), hi = bf_hi(s0[i]) * bf_hi(k3[i]);
        lo = fmaf(bf_lo(s1[i]), bf_lo(k2[i]), lo); hi = fmaf(bf_hi(s1[i]), bf_hi(k2[i]), hi);
        lo = fmaf(bf_lo(s2[i]), bf_lo(k1[i]), lo); hi = fmaf(bf_hi(s2[i]), bf_hi(k1[i]), hi);
        lo = fmaf(bf_lo(s3[i]), bf_lo(k0[i]), lo); hi = fmaf(bf_hi(s3[i]), bf_hi(k0[i]), hi);
        o[i] = pack_bf2(lo + bf_lo(bb[i]), hi + bf_hi(bb[i]));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float acc = __fmul_rn(__uint_as_float(s0[i]), __uint_as_float(k3[i]));
        acc = __fadd_rn(acc, __fmul_rn(__uint_as_float(s1[i]), __uint_as_float(k2[i])));
        acc = __fadd_rn(acc, __fmul_rn(__uint_as_float(s2[i]), __uint_as_float(k1[i])));
        acc = __fadd_rn(acc, __fmul_rn(__uint_as_float(s3[i]), __uint_as_float(k0[i])));
        o[i] = __float_as_uint(__fadd_rn(acc, __uint_as_float(bb[i])));
      }
    }
    // CG_CONV_ST: 0 = streaming (evict-first) store, 1 = default policy, 2 = L2::evict_last.  The
    // consumer (RG-LRU kernel) re-reads the 84 MB of a config-2 output right away: kept in the
    // 126 MB L2 they need no DRAM read and their write-back leaves this kernel's critical path
#if CG_CONV_ST == 1
    *reinterpret_cast<uint4*>(yb + (size_t)t * p.E) = make_uint4(o[0], o[1], o[2], o[3]);
#elif CG_CONV_ST == 2
    {
      uint64_t pol;
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
      asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;"
                   :: "l"(yb + (size_t)t * p.E), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "l"(pol) : "memory");
    }
#else
    stg_stream(yb + (size_t)t * p.E, make_uint4(o[0], o[1], o[2], o[3]));
#endif
  }

  // new cache = last 3 input rows, left zero padded (:542-543); written by the
  // thread whose tile holds the final step.
  if (p.cache_out != nullptr && t0 + LC >= p.T) {
    IO* cb = reinterpret_cast<IO*>(p.cache_out) + (size_t)b * 3 * p.E + ch0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ti = p.T - 3 + r;
      const uint4 v = ti >= 0 ? ldg_stream(xb + (size_t)ti * p.E) : zero;
      *reinterpret_cast<uint4*>(cb + (size_t)r * p.E) = v;
    }
  }
}

// Grid-mapped launch: blockIdx.x walks the channel tiles of a row first, so
// concurrently running blocks sweep whole rows (DRAM page locality).
template <typename IO, bool EMUL, int LC>
__global__ void __launch_bounds__(128)
conv1d_w4_kernel(const ConvParams p) {
  // programmatic dependent launch: once every block of this grid has started, a
  // PDL consumer (the RG-LRU prologue, then the fused kernel's set-up) may begin;
  // consumers order themselves after this grid's stores with griddepcontrol.wait
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  conv1d_w4_tile<IO, EMUL, LC>(p, blockIdx.x, blockIdx.y, blockIdx.z);
}

// ---------------------------------------------------------------------------
// Backward of the W == 4 prefill (what autograd derives from layers.py:484-546;
// training path).  With m_s(t) the document mask
// of tap s at output step t (the forward's m1..m3):
//   dx[u]    = sum_s w[3-s] * gy[u+s] * m_s(u+s)
//   dw[3-s]  = sum_{b,t} gy[t] * x[t-s] * m_s(t)        db = sum_{b,t} gy[t]
// Same tiling as the forward: thread = 16-byte channel vector x LC steps.  dx is
// accumulated in the tensor dtype in autograd's order (bit-exact with the
// reference); the parameter gradients are summed in fp32 and reduced
// in a fixed order (registers -> shared memory over the block's 16 time slots
// -> one partial row per block -> conv1d_bwd_reduce_kernel): bit-reproducible.
// ---------------------------------------------------------------------------
constexpr int kConvBwdSpan = 8;

struct ConvBwdParams {
  const void* gy;     // [B,T,E]
  const void* x;      // [B,T,E] forward input
  const void* w;      // [4,E]
  const void* seg;
  long long seg_bstride;
  int seg_is_i64;
  void* dx;           // [B,T,E]
  float* partial;     // [B * tblocks][5][E] fp32: dw[0..3], db per block
  int B, T, E;
  int mask_mode;
};

template <typename IO>
__device__ __forceinline__ void widen_vec(const uint4& v, float (&f)[IoVec<IO>::V]) {
  if constexpr (IoVec<IO>::kBf16) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[2 * i] = bf_lo(w[i]); f[2 * i + 1] = bf_hi(w[i]); }
  } else {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
    f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
  }
}

// (forcing more resident blocks with a minBlocks bound was measured: 4 -> 79 us,
// 5 -> 156 us against 75 us unconstrained at 143 registers)
template <typename IO, int LC>
__global__ void __launch_bounds__(128)
conv1d_w4_bwd_kernel(const ConvBwdParams p) {
  constexpr int V = IoVec<IO>::V;
  constexpr bool BF = IoVec<IO>::kBf16;
  constexpr int EC = kCvl * V;
  __shared__ float red[16][kCvl][5 * V + 1];
  const int cv = threadIdx.x & 7;
  const int slot = threadIdx.x >> 3;
  const int b = blockIdx.z;
  const int ch0 = blockIdx.x * EC + cv * V;
  float acc[5][V];
#pragma unroll
  for (int k = 0; k < 5; ++k)
#pragma unroll
    for (int i = 0; i < V; ++i) acc[k][i] = 0.0f;

  // kConvBwdSpan consecutive time blocks (16 slots x LC steps each) per CTA: the
  // parameter-gradient partials stay in registers across them
#pragma unroll 1
  for (int q = 0; q < kConvBwdSpan; ++q) {
    const int tslot = (blockIdx.y * kConvBwdSpan + q) * 16 + slot;
    const int t0 = tslot * LC;
    if (ch0 >= p.E || t0 >= p.T) break;
    const IO* gb = reinterpret_cast<const IO*>(p.gy) + (size_t)b * p.T * p.E + ch0;
    const IO* xb = reinterpret_cast<const IO*>(p.x) + (size_t)b * p.T * p.E + ch0;
    IO* dxb = reinterpret_cast<IO*>(p.dx) + (size_t)b * p.T * p.E + ch0;
    const long long seg0 = (long long)b * p.seg_bstride;
    const uint4 zero = make_uint4(0, 0, 0, 0);
    uint4 gr[LC + 3], xr[LC + 3];   // gr[r] = gy[t0 + r], xr[r] = x[t0 - 3 + r]
#pragma unroll
    for (int r = 0; r < LC + 3; ++r) {
      const int tg = t0 + r, tx = t0 - 3 + r;
      gr[r] = tg < p.T ? ldg_stream(gb + (size_t)tg * p.E) : zero;
      xr[r] = (tx >= 0 && tx < p.T) ? ldg_stream(xb + (size_t)tx * p.E) : zero;
    }
    // nb bit r = (segment_pos[t0 - 2 + r] != 0), r < LC + 5 (steps t0-2 .. t0+LC+2)
    unsigned nb = 0;
#pragma unroll
    for (int r = 0; r < LC + 5; ++r) {
      const int t = t0 - 2 + r;
      if (t >= 0 && t < p.T)
        nb |= (load_seg(p.seg, p.seg_is_i64 != 0, seg0 + t) != 0 ? 1u : 0u) << r;
    }
    // masks of output step t (as the forward): bit 0 seg[t-2], bit 1 seg[t-1], bit 2 seg[t]
    auto masks = [&](int t, bool& m1, bool& m2, bool& m3) {
      const unsigned w3 = nb >> (t - t0);
      if (p.mask_mode == 0) { m1 = true; m2 = true; m3 = (w3 & 1u) != 0; }
      else { m1 = (w3 & 4u) != 0; m2 = (w3 & 6u) == 6u; m3 = (w3 & 7u) == 7u; }
    };
    // taps as loaded (packed bf16x2 pairs or fp32 words): wq[k][i]
    uint32_t wq[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const IO*>(p.w) + (size_t)k * p.E + ch0);
      wq[k][0] = v.x; wq[k][1] = v.y; wq[k][2] = v.z; wq[k][3] = v.w;
    }
    // one product / one sum in the tensor dtype, rounded as the eager op is
    auto mul_io = [](uint32_t u, uint32_t v) -> uint32_t {
      if constexpr (BF) return bf2_mul(u, v);
      else return __float_as_uint(__fmul_rn(__uint_as_float(u), __uint_as_float(v)));
    };
    auto add_io = [](uint32_t u, uint32_t v) -> uint32_t {
      if constexpr (BF) return bf2_add(u, v);
      else return __float_as_uint(__fadd_rn(__uint_as_float(u), __uint_as_float(v)));
    };

#pragma unroll
    for (int j = 0; j < LC; ++j) {
      const int t = t0 + j;
      if (t >= p.T) break;
      const uint32_t g0[4] = {gr[j].x, gr[j].y, gr[j].z, gr[j].w};
      const uint32_t g1[4] = {gr[j + 1].x, gr[j + 1].y, gr[j + 1].z, gr[j + 1].w};
      const uint32_t g2[4] = {gr[j + 2].x, gr[j + 2].y, gr[j + 2].z, gr[j + 2].w};
      const uint32_t g3[4] = {gr[j + 3].x, gr[j + 3].y, gr[j + 3].z, gr[j + 3].w};
      // ---- dx[t]: taps of the output steps t .. t+3 that read x[t].  Autograd
      // accumulates the four contributions in the tensor dtype, last tap first
      // (tap 3, 2, 1, 0), each product rounded: reproduced exactly.
      bool a1, a2, a3, b1, b2, b3, c1, c2, c3;
      masks(t + 1, a1, a2, a3); masks(t + 2, b1, b2, b3); masks(t + 3, c1, c2, c3);
      uint32_t o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t v = c3 ? mul_io(g3[i], wq[0][i]) : 0u;          // gy rows beyond T were loaded as zero
        v = add_io(v, b2 ? mul_io(g2[i], wq[1][i]) : 0u);
        v = add_io(v, a1 ? mul_io(g1[i], wq[2][i]) : 0u);
        o[i] = add_io(v, mul_io(g0[i], wq[3][i]));
      }
      stg_stream(dxb + (size_t)t * p.E, make_uint4(o[0], o[1], o[2], o[3]));
      // ---- parameter gradients of output step t: products rounded to the
      // tensor dtype (the eager mul of autograd), summed in fp32
      bool m1, m2, m3;
      masks(t, m1, m2, m3);
      const uint4 zr = make_uint4(0, 0, 0, 0);
      const uint4 r0 = xr[j + 3], r1 = m1 ? xr[j + 2] : zr, r2 = m2 ? xr[j + 1] : zr, r3 = m3 ? xr[j] : zr;
      const uint32_t x0[4] = {r0.x, r0.y, r0.z, r0.w}, x1[4] = {r1.x, r1.y, r1.z, r1.w};
      const uint32_t x2[4] = {r2.x, r2.y, r2.z, r2.w}, x3[4] = {r3.x, r3.y, r3.z, r3.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t q3 = mul_io(g0[i], x0[i]), q2 = mul_io(g0[i], x1[i]);
        const uint32_t q1 = mul_io(g0[i], x2[i]), q0 = mul_io(g0[i], x3[i]);
        if constexpr (BF) {
          acc[3][2 * i] += bf_lo(q3); acc[3][2 * i + 1] += bf_hi(q3);
          acc[2][2 * i] += bf_lo(q2); acc[2][2 * i + 1] += bf_hi(q2);
          acc[1][2 * i] += bf_lo(q1); acc[1][2 * i + 1] += bf_hi(q1);
          acc[0][2 * i] += bf_lo(q0); acc[0][2 * i + 1] += bf_hi(q0);
          acc[4][2 * i] += bf_lo(g0[i]); acc[4][2 * i + 1] += bf_hi(g0[i]);
        } else {
          acc[3][i] += __uint_as_float(q3); acc[2][i] += __uint_as_float(q2);
          acc[1][i] += __uint_as_float(q1); acc[0][i] += __uint_as_float(q0);
          acc[4][i] += __uint_as_float(g0[i]);
        }
      }
    }
  }
  // ---- block reduction over the 16 time slots, fixed order
#pragma unroll
  for (int k = 0; k < 5; ++k)
#pragma unroll
    for (int i = 0; i < V; ++i) red[slot][cv][k * V + i] = acc[k][i];
  __syncthreads();
  float* out = p.partial + ((size_t)(b * gridDim.y + blockIdx.y) * 5) * p.E;
  for (int idx = threadIdx.x; idx < kCvl * 5 * V; idx += blockDim.x) {
    const int c = idx / (5 * V), r = idx - c * (5 * V);
    const int k = r / V, i = r - k * V;
    const int ch = blockIdx.x * EC + c * V + i;
    if (ch >= p.E) continue;
    float sum = 0.0f;
#pragma unroll
    for (int sl = 0; sl < 16; ++sl) sum += red[sl][c][r];
    out[(size_t)k * p.E + ch] = sum;
  }
}

// dw[k][e] / db[e] = sum of the per-block partials, in block order.
template <typename IO>
__global__ void conv1d_bwd_reduce_kernel(const float* __restrict__ partial, int nparts, int E,
                                         void* __restrict__ dw, void* __restrict__ db) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 5 * E) return;
  float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;   // four interleaved chains, combined in a fixed order
  int q = 0;
  for (; q + 3 < nparts; q += 4) {
    s0 += partial[(size_t)q * 5 * E + idx];
    s1 += partial[(size_t)(q + 1) * 5 * E + idx];
    s2 += partial[(size_t)(q + 2) * 5 * E + idx];
    s3 += partial[(size_t)(q + 3) * 5 * E + idx];
  }
  for (; q < nparts; ++q) s0 += partial[(size_t)q * 5 * E + idx];
  const float sum = (s0 + s1) + (s2 + s3);
  if (idx < 4 * E) store_io<IO>(dw, idx, sum);
  else store_io<IO>(db, idx - 4 * E, sum);
}

// ---------------------------------------------------------------------------
// Generic temporal width (W != 4): one thread per output element.  Slow path,
// kept for API completeness (the reference's tests also run W = 8).
// Reproduces the accumulated in-place masking of the returned cache
// (quirk D2, see oracle/rglru_oracle.c).
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool tap_mask_generic(const ConvParams& p, long long seg0, int ti,
                                                 int shift, int mask_mode) {
  const int hi = mask_mode == 0 ? shift - 2 : shift;
  for (int j = 1; j <= hi; ++j)
    if (load_seg(p.seg, p.seg_is_i64 != 0, seg0 + ti + j) == 0) return false;
  return true;
}

template <typename IO, bool EMUL>
__global__ void conv1d_generic_kernel(const ConvParams p) {
  constexpr bool BF = IoVec<IO>::kBf16;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y;
  const int b = blockIdx.z;
  if (e >= p.E) return;
  const long long seg0 = (long long)b * p.seg_bstride;
  const int taps = p.W < p.T ? p.W : p.T;
  float acc = 0.0f;
  for (int s = 0; s < taps; ++s) {
    const int ti = t - s;
    float term = 0.0f;
    if (ti >= 0) {
      float xv = load_io<IO>(p.x, ((size_t)b * p.T + ti) * p.E + e);
      if (!tap_mask_generic(p, seg0, ti, s, p.mask_mode)) xv = 0.0f;
      term = __fmul_rn(xv, load_io<IO>(p.w, (size_t)(p.W - 1 - s) * p.E + e));
      if (BF && EMUL) term = round_bf(term);
    }
    acc = s == 0 ? term : __fadd_rn(acc, term);
    if (BF && EMUL) acc = round_bf(acc);
  }
  acc = __fadd_rn(acc, load_io<IO>(p.bias, e));
  store_io<IO>(p.y, ((size_t)b * p.T + t) * p.E + e, acc);

  if (p.cache_out != nullptr && t == p.T - 1) {
    for (int r = 0; r < p.W - 1; ++r) {
      const int ti = p.T - (p.W - 1) + r;
      float xv = 0.0f;
      if (ti >= 0) {
        xv = load_io<IO>(p.x, ((size_t)b * p.T + ti) * p.E + e);
        if (p.mask_mode == 0) {
          int smax = taps - 1 < p.T - 1 - ti ? taps - 1 : p.T - 1 - ti;
          if (smax >= 3 && !tap_mask_generic(p, seg0, ti, smax, 0)) xv = 0.0f;
        }
      }
      store_io<IO>(p.cache_out, ((size_t)b * (p.W - 1) + r) * p.E + e, xv);
    }
  }
}

// ---------------------------------------------------------------------------
// Decode step (T == 1, cache given, no mask): one thread per (b, channel).
// cache_out may alias cache_in (each thread reads its column before writing).
// ---------------------------------------------------------------------------
template <typename IO, bool EMUL>
__global__ void conv1d_decode_kernel(const ConvParams p) {
  constexpr bool BF = IoVec<IO>::kBf16;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (e >= p.E) return;
  const int W = p.W;
  auto cache_at = [&](int k) -> float {   // cache.type(x.dtype), :567
    const size_t i = ((size_t)b * (W - 1) + k) * p.E + e;
    float v = p.cache_is_bf16
                  ? __uint_as_float((uint32_t)reinterpret_cast<const uint16_t*>(p.cache_in)[i] << 16)
                  : reinterpret_cast<const float*>(p.cache_in)[i];
    if (BF) v = round_bf(v);
    return v;
  };
  const float xnew = load_io<IO>(p.x, (size_t)b * p.E + e);
  float acc = 0.0f;
  for (int s = 0; s < W; ++s) {
    const int k = W - 1 - s;
    const float xv = k == W - 1 ? xnew : cache_at(k);
    float term = __fmul_rn(xv, load_io<IO>(p.w, (size_t)k * p.E + e));
    if (BF && EMUL) term = round_bf(term);
    acc = s == 0 ? term : __fadd_rn(acc, term);
    if (BF && EMUL) acc = round_bf(acc);
  }
  acc = __fadd_rn(acc, load_io<IO>(p.bias, e));
  store_io<IO>(p.y, (size_t)b * p.E + e, acc);
  if (p.cache_out != nullptr) {
    // new_cache[k-1] = xcat[k] (:542): read all surviving rows first so that
    // cache_out may alias cache_in, then write
    float keep[16];
    for (int k = 1; k < W; ++k) keep[k - 1] = k == W - 1 ? xnew : cache_at(k);
    for (int k = 1; k < W; ++k) {
      const size_t i = ((size_t)b * (W - 1) + (k - 1)) * p.E + e;
      if (p.cache_is_bf16) reinterpret_cast<uint16_t*>(p.cache_out)[i] = (uint16_t)(pack_bf2(keep[k - 1], keep[k - 1]) & 0xffffu);
      else reinterpret_cast<float*>(p.cache_out)[i] = keep[k - 1];
    }
  }
}

}  // namespace cg
