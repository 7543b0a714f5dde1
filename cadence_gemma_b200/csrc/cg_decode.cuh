// Fused decode step of the recurrent hot path for sm_100a (T == 1, caches
// given): temporal Conv1D step + cache roll, both block-diagonal gate GEMVs, gate
// math and the one-step recurrence h = a*h0 + x~ in ONE launch.
//
// Replaces, for a decode step, reference recurrentgemma/torch/layers.py:478-483 +
// :542 (Conv1D with cache), :133-142 (BlockDiagonalLinear x 2), :345-365 (gates)
// and :175-182 (rnn_scan's sampling branch) -- the three launches (conv step,
// cuBLAS batched GEMV, gate step) the unfused path needs per block and token.
// SURVEY.md section 8(b) `cg_recurrent_decode_step`, 8(f) row F3.
//
// Decomposition: CTA = (head, 64 output channels of the head, BT batch rows),
// 256 threads.  Phase 1: the CTA evaluates the convolution step for all `bw`
// input channels of its head and its rows into shared memory (every rounding
// point of the reference, as conv1d_decode_kernel) and -- the channel-slice-0
// CTA only -- writes the rolled cache.  Phase 2: thread (channel pair, K octant)
// accumulates both gates' dot products over its eighth of K for all BT rows in
// fp32 (weights read coalesced, 128 B per warp and k; activations broadcast from
// shared memory).  Phase 3: the eight octants are added in a fixed order, the sum
// is rounded to bf16 (the GEMV output the reference materialises), and thread
// (channel j, q) runs the gate math on the bf16x2 pair of rows (2q, 2q+1).
// The CTA's [2 gates x bw x 64] weight slice is fetched with cp.async at kernel
// start (one L2 round trip, under phase 1).
#pragma once

#include "cg_common.cuh"
#include "cg_scan.cuh"

namespace cg {

constexpr int kDecBT = 8;       // batch rows per CTA
constexpr int kDecMaxBw = 256;  // head width limit (shared memory staging)

struct DecodeParams {
  const uint16_t* x;         // [B,1,E]
  const uint16_t* conv_w;    // [4,E]
  const uint16_t* conv_b;    // [E]
  const void* cache_in;      // [B,3,E] bf16 or fp32
  int cache_is_bf16;
  const uint16_t* wx;        // [H,bw,bw] input gate (y = x @ w[h])
  const uint16_t* wa;        // [H,bw,bw] a gate
  const uint16_t* bias_x;    // [E] or null
  const uint16_t* bias_a;
  const uint16_t* a_param;   // [E]
  const void* seg;           // [B,1]
  int seg_is_i64;
  long long seg_bstride;
  const float* h0;           // [B,E] or null
  const uint16_t* gate_mul;  // [B,1,E] or null: y <- round(y * gate_mul) (modules.py:651)
  uint16_t* y;               // [B,1,E]
  void* cache_out;           // [B,3,E] in the cache dtype, or null; must not alias cache_in
  float* last_h;             // [B,E] or null
  int B, E, H, bw;
};

template <bool FAST>
__global__ void __launch_bounds__(256)
recurrent_decode_kernel(const DecodeParams p) {
  __shared__ float s_xc[kDecBT][kDecMaxBw];            // conv output (bf16 values) of the head
  extern __shared__ __align__(16) unsigned char s_dyn[];   // [2 gates][bw][64 channels] bf16 weight slice
  // per-octant partial sums [8][BT][2][64] fp32 (32 KB): they take the place of
  // the weight slice once phase 2 has consumed it (>= 32 KB for bw >= 128; the
  // launch reserves max(weights, 32 KB))
  float (*s_red)[kDecBT][2][64] = reinterpret_cast<float (*)[kDecBT][2][64]>(s_dyn);
  const int bw = p.bw;
  const int slices = bw / 64;
  const int head = blockIdx.x / slices, slice = blockIdx.x - head * slices;
  const int b0 = blockIdx.y * kDecBT;
  const int tid = threadIdx.x;

  // ---- phase 0: the CTA's slice of both gate matrices streams into shared memory
  // (cp.async, all requests in flight at once) while phase 1 runs
  {
    const int chunks = 2 * bw * 8;                     // 16-byte chunks: [gate][k][8 per 128 B row]
    for (int id = tid; id < chunks; id += 256) {
      const int c8 = id & 7, k = (id >> 3) % bw, g = id / (8 * bw);
      const uint16_t* src = (g == 0 ? p.wx : p.wa) + (size_t)head * bw * bw + (size_t)k * bw + slice * 64 + c8 * 8;
      cp_async16(s_dyn + (size_t)id * 16, src);
    }
    cp_async_commit();
  }

  // ---- phase 1: convolution step (layers.py:478-483, accumulation order :533-536).
  // Thread = one channel of the head (bw <= 256), all BT rows: every load of the
  // step is issued before the first use.
  if (tid < bw) {
    const int ch = head * bw + tid;
    float wv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) wv[k] = __uint_as_float((uint32_t)p.conv_w[(size_t)k * p.E + ch] << 16);
    const float bias = __uint_as_float((uint32_t)p.conv_b[ch] << 16);
    float xin[kDecBT][4];                              // [row][cache0, cache1, cache2, x]
#pragma unroll
    for (int r = 0; r < kDecBT; ++r) {
      const int b = b0 + r;
#pragma unroll
      for (int k = 0; k < 4; ++k) xin[r][k] = 0.0f;
      if (b < p.B) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const size_t i = ((size_t)b * 3 + k) * p.E + ch;
          xin[r][k] = p.cache_is_bf16
                          ? __uint_as_float((uint32_t)reinterpret_cast<const uint16_t*>(p.cache_in)[i] << 16)
                          : reinterpret_cast<const float*>(p.cache_in)[i];
        }
        xin[r][3] = __uint_as_float((uint32_t)p.x[(size_t)b * p.E + ch] << 16);
      }
    }
#pragma unroll
    for (int r = 0; r < kDecBT; ++r) {
      const int b = b0 + r;
      const float c0v = round_bf(xin[r][0]), c1v = round_bf(xin[r][1]), c2v = round_bf(xin[r][2]);   // cache.type(x.dtype), :567
      const float xnew = xin[r][3];
      float acc = round_bf(__fmul_rn(xnew, wv[3]));    // shift s multiplies w[3 - s]
      acc = round_bf(__fadd_rn(acc, round_bf(__fmul_rn(c2v, wv[2]))));
      acc = round_bf(__fadd_rn(acc, round_bf(__fmul_rn(c1v, wv[1]))));
      acc = round_bf(__fadd_rn(acc, round_bf(__fmul_rn(c0v, wv[0]))));
      acc = round_bf(__fadd_rn(acc, bias));
      s_xc[r][tid] = b < p.B ? acc : 0.0f;
      if (slice == 0 && p.cache_out != nullptr && b < p.B) {   // new cache = [c1, c2, x] (:542)
        const float keep[3] = {c1v, c2v, xnew};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const size_t i = ((size_t)b * 3 + k) * p.E + ch;
          if (p.cache_is_bf16)
            reinterpret_cast<uint16_t*>(p.cache_out)[i] = (uint16_t)(pack_bf2(keep[k], keep[k]) & 0xffffu);
          else
            reinterpret_cast<float*>(p.cache_out)[i] = keep[k];
        }
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- phase 2: thread (channel pair jp, K octant o) accumulates both gates' dot
  // products over its eighth of K for two channels and all BT rows (fp32).  A
  // warp reads 128 contiguous bytes of weights per k and gate.
  {
    const int jp = tid & 31, o = tid >> 5;
    const int kc = bw / 8;
    float ax[kDecBT][2], aa[kDecBT][2];
#pragma unroll
    for (int r = 0; r < kDecBT; ++r) { ax[r][0] = ax[r][1] = aa[r][0] = aa[r][1] = 0.0f; }
    const uint32_t* wxp = reinterpret_cast<const uint32_t*>(s_dyn) + (size_t)(o * kc) * 32 + jp;
    const uint32_t* wap = wxp + (size_t)bw * 32;       // second gate: bw rows of 32 words further
#pragma unroll 8
    for (int k = 0; k < kc; ++k) {
      const uint32_t wx2 = wxp[k * 32];
      const uint32_t wa2 = wap[k * 32];
      const float wx0 = bf_lo(wx2), wx1 = bf_hi(wx2), wa0 = bf_lo(wa2), wa1 = bf_hi(wa2);
#pragma unroll
      for (int r = 0; r < kDecBT; ++r) {
        const float xv = s_xc[r][o * kc + k];
        ax[r][0] = fmaf(xv, wx0, ax[r][0]); ax[r][1] = fmaf(xv, wx1, ax[r][1]);
        aa[r][0] = fmaf(xv, wa0, aa[r][0]); aa[r][1] = fmaf(xv, wa1, aa[r][1]);
      }
    }
    __syncthreads();                                   // every warp is done with the weights
#pragma unroll
    for (int r = 0; r < kDecBT; ++r) {
      *reinterpret_cast<float2*>(&s_red[o][r][0][2 * jp]) = make_float2(ax[r][0], ax[r][1]);
      *reinterpret_cast<float2*>(&s_red[o][r][1][2 * jp]) = make_float2(aa[r][0], aa[r][1]);
    }
  }
  __syncthreads();

  const int j = tid & 63, q = tid >> 6;
  const int chj = slice * 64 + j;                      // output channel inside the head
  // ---- phase 3: thread (j, q) finishes rows (2q, 2q+1) of channel `ch`
  const int ch = head * bw + chj;
  const int r0 = 2 * q, b_lo = b0 + r0, b_hi = b_lo + 1;
  if (b_lo >= p.B) return;
  float sx[2], sa[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {                        // octants added in a fixed order
    sx[h] = s_red[0][r0 + h][0][j];
    sa[h] = s_red[0][r0 + h][1][j];
#pragma unroll
    for (int o = 1; o < 8; ++o) { sx[h] += s_red[o][r0 + h][0][j]; sa[h] += s_red[o][r0 + h][1][j]; }
  }
  const uint32_t gxr = pack_bf2(sx[0], sx[1]);         // the GEMV output the reference materialises in bf16
  const uint32_t gar = pack_bf2(sa[0], sa[1]);
  const uint32_t xc2 = pack_bf2(s_xc[r0][chj], s_xc[r0 + 1][chj]);
  uint32_t bx2 = 0, ba2 = 0;
  if (p.bias_x != nullptr) { const uint32_t v = p.bias_x[ch]; bx2 = v | (v << 16); }
  if (p.bias_a != nullptr) { const uint32_t v = p.bias_a[ch]; ba2 = v | (v << 16); }
  const float ap = __uint_as_float((uint32_t)p.a_param[ch] << 16);
  const float sp8 = -8.0f * round_bf(softplus_f(ap));  // softplus rounded to bf16 first (:352)
  const uint32_t sp2 = pack_bf2(sp8, sp8);
  uint32_t a2, n2;
  gate_pair_emul<FAST, false>(xc2, gxr, gar, bx2, ba2, sp2, a2, n2);
  const bool rs_lo = seg_is_zero(p.seg, p.seg_is_i64 != 0, (long long)b_lo * p.seg_bstride);
  const bool rs_hi = b_hi < p.B && seg_is_zero(p.seg, p.seg_is_i64 != 0, (long long)b_hi * p.seg_bstride);
  if (rs_lo || rs_hi) {                                // a document start at a decode step (:173, :364)
    uint32_t az, nr;
    gate_pair_emul<FAST, true>(xc2, gxr, gar, bx2, ba2, sp2, az, nr);
    if (rs_lo) { a2 &= 0xffff0000u; n2 = (n2 & 0xffff0000u) | (nr & 0x0000ffffu); }
    if (rs_hi) { a2 &= 0x0000ffffu; n2 = (n2 & 0x0000ffffu) | (nr & 0xffff0000u); }
  }
  // rnn_scan, T == 1 (:175-182): y = a*h0 + x~ (separate fp32 mul and add), or x~ without h0
  float hl, hh;
  if (p.h0 == nullptr) {
    hl = bf_lo(n2); hh = bf_hi(n2);
  } else {
    hl = __fadd_rn(__fmul_rn(bf_lo(a2), p.h0[(size_t)b_lo * p.E + ch]), bf_lo(n2));
    hh = b_hi < p.B ? __fadd_rn(__fmul_rn(bf_hi(a2), p.h0[(size_t)b_hi * p.E + ch]), bf_hi(n2)) : 0.0f;
  }
  uint32_t o = pack_bf2(hl, hh);
  if (p.gate_mul != nullptr) {
    const uint32_t g = p.gate_mul[(size_t)b_lo * p.E + ch] |
                       (b_hi < p.B ? (uint32_t)p.gate_mul[(size_t)b_hi * p.E + ch] << 16 : 0u);
    o = bf2_mul(o, g);
  }
  p.y[(size_t)b_lo * p.E + ch] = (uint16_t)(o & 0xffffu);
  if (p.last_h != nullptr) p.last_h[(size_t)b_lo * p.E + ch] = hl;
  if (b_hi < p.B) {
    p.y[(size_t)b_hi * p.E + ch] = (uint16_t)(o >> 16);
    if (p.last_h != nullptr) p.last_h[(size_t)b_hi * p.E + ch] = hh;
  }
}

}  // namespace cg
