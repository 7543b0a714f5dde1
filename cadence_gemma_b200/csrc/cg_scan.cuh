// Chunked two-level linear-recurrence scan for sm_100a, fused with the RG-LRU
// gate math.  Replaces reference recurrentgemma/torch/layers.py:345-375
// (RGLRU.forward after the gate GEMMs) and :146-199 (rnn_scan).
//
// Decomposition (one WARP per work item, no shared memory, no block barrier):
//   item   = (batch row b, channel tile of EC = 8*V channels, chunk of TC steps)
//   lane   = (seg = lane / 8, cv = lane % 8):  8 lanes x V channels cover one
//            contiguous 128-byte row (V = 8 bf16 / 4 fp32 -> 16-byte vectors),
//            the 4 lane groups take 4 consecutive time segments of L steps
//            => every warp load/store instruction moves 4 full 128 B lines.
//   pass 1 : each lane streams its L steps, evaluates the gates in registers
//            and keeps (a_t, x~_t) (packed bf16x2 when they are bf16-exact),
//            accumulating the segment transform h -> P*h + H.
//   carry  : warp-shuffle scan of (P,H) over the 4 segments; chunks of one
//            column are chained through global memory with a decoupled
//            look-back (flag 1 = aggregate published, 2 = inclusive state
//            published).  The carry is always folded left-to-right over the
//            chunk aggregates, so results do not depend on timing.
//   pass 2 : replay h = a_t*h + x~_t from the true carry-in, store y.
// Items are handed out by an atomic ticket in time-major order, so every
// item's predecessors are already running: the look-back cannot deadlock.
#pragma once

#include "cg_common.cuh"

namespace cg {

constexpr int kSegs = 4;   // time segments per warp (lane / 8)
constexpr int kCvl = 8;    // lanes across channels (lane % 8)

struct ScanParams {
  // RG-LRU inputs (KIND 0)
  const void* x;        // [B,T,E]
  const void* gemm_x;   // [B,T,*] row stride gate_ld
  const void* gemm_a;
  const void* bias_x;   // [E] or null
  const void* bias_a;
  const float* neg8sp;  // [E] fp32: -8*softplus(a_param) (rounded per mode)
  const void* seg;      // segment positions
  long long seg_bstride;
  int seg_is_i64;
  long long gate_ld;
  // plain scan inputs (KIND 1): x above, plus
  const void* a;                 // [B,T,E]
  const unsigned char* reset;    // [B,T]
  // state
  const float* h0;   // [B,E] or null
  void* y;           // [B,T,E]
  float* last_h;     // [B,E] or null
  // scratch
  int* counter;      // ticket
  int* flags;        // [nitems]
  float* agg_p;      // [nitems][EC]
  float* agg_h;
  float* pref;       // [nitems][EC]
  int B, T, E;
  int ncols;         // B * ceil(E / EC)
  int ctiles;        // ceil(E / EC)
  int nchunks;
  int nitems;
};

template <typename IO> struct IoVec;
template <> struct IoVec<uint16_t> { static constexpr int V = 8; static constexpr bool kBf16 = true; };
template <> struct IoVec<float> { static constexpr int V = 4; static constexpr bool kBf16 = false; };

// ---- RG-LRU gates on one bf16x2 pair, every eager rounding point reproduced.
// layers.py:348-365 + :173; see SURVEY.md section 7 "bf16 rounding points".
template <bool FAST>
__device__ __forceinline__ void gate_pair_emul(uint32_t xc, uint32_t gxr, uint32_t gar,
                                               uint32_t bx, uint32_t ba, uint32_t sp8,
                                               bool reset, uint32_t& a_out, uint32_t& nx_out) {
  const uint32_t px = bf2_add(gxr, bx);          // r(gemm + b)            :139
  const uint32_t pa = bf2_add(gar, ba);
  float sx0, sx1, sa0, sa1;
  if constexpr (FAST) {                          // pair each x gate with an a gate
    sigmoid2<true>(bf_lo(px), bf_lo(pa), sx0, sa0);
    sigmoid2<true>(bf_hi(px), bf_hi(pa), sx1, sa1);
  } else {
    sigmoid2<false>(bf_lo(px), bf_hi(px), sx0, sx1);
    sigmoid2<false>(bf_lo(pa), bf_hi(pa), sa0, sa1);
  }
  const uint32_t gx = pack_bf2(sx0, sx1);        // r(sigmoid)             :348
  const uint32_t ga = pack_bf2(sa0, sa1);        //                        :349
  const uint32_t la = bf2_mul(ga, sp8);          // r(r(-8 ga) * sp)       :352
  const float l0 = bf_lo(la), l1 = bf_hi(la);
  float e0, e1, q0, q1;
  if constexpr (FAST) {
    e0 = ex2_approx(l0 * kLog2e); e1 = ex2_approx(l1 * kLog2e);
    q0 = e0 * e0; q1 = e1 * e1;
  } else {
    e0 = expf(l0); e1 = expf(l1);                //                        :353
    q0 = expf(2.0f * l0); q1 = expf(2.0f * l1);  // 2*log_a is exact       :354
  }
  uint32_t av = pack_bf2(e0, e1);
  const uint32_t om = bf2_sub(kOne2, pack_bf2(q0, q1));   // r(1 - a^2)    :361
  uint32_t mu = pack_bf2(sqrt_f<FAST>(bf_lo(om)), sqrt_f<FAST>(bf_hi(om)));
  if (reset) { mu = kOne2; av = 0u; }            //                  :364, :173
  nx_out = bf2_mul(bf2_mul(xc, gx), mu);         // r(r(x*gx)*mu)    :357, :365
  a_out = av;
}

// ---- same, fp32 kept in registers (bf16 or fp32 inputs already widened).
// For fp32 tensors this IS the reference op order (each op rounds to fp32).
template <bool FAST>
__device__ __forceinline__ void gate_f32(float xc, float gxr, float gar, float bx, float ba,
                                         float sp8, bool reset, float& a_out, float& nx_out) {
  const float px = gxr + bx, pa = gar + ba;
  float gx, ga;
  sigmoid2<FAST>(px, pa, gx, ga);
  const float la = ga * sp8;
  const float e = exp_f<FAST>(la);
  float q;
  if constexpr (FAST) q = e * e; else q = expf(2.0f * la);
  float mu = sqrt_f<FAST>(1.0f - q);
  float av = e;
  if (reset) { mu = 1.0f; av = 0.0f; }
  nx_out = __fmul_rn(__fmul_rn(xc, gx), mu);
  a_out = av;
}

// KIND: 0 = RG-LRU (gates fused), 1 = plain rnn_scan(x, a, reset, h0).
// ARITH: CG_ARITH_* bits 0 (fp32 in registers) and 1 (fast math).
template <typename IO, int KIND, int ARITH, int L, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
scan_kernel(const ScanParams p) {
  constexpr int V = IoVec<IO>::V;
  constexpr bool BF = IoVec<IO>::kBf16;
  constexpr bool FAST = (ARITH & 2) != 0;
  // (a, x~) are exactly bf16 -> keep them packed between the two passes
  constexpr bool PACKED = BF && (KIND == 1 || (ARITH & 1) == 0);
  constexpr int EC = kCvl * V;
  constexpr int TC = kSegs * L;
  constexpr int NV = PACKED ? V / 2 : V;   // registers per stored vector

  const int lane = threadIdx.x & 31;
  const int seg_id = lane >> 3;
  const int cv = lane & 7;

  int item = 0;
  if (lane == 0) item = atomicAdd(p.counter, 1);
  item = __shfl_sync(0xffffffffu, item, 0);
  if (item >= p.nitems) return;
  const int chunk = item / p.ncols;          // time-major ticket order
  const int col = item - chunk * p.ncols;
  const int b = col / p.ctiles;
  const int ch0 = (col - b * p.ctiles) * EC + cv * V;
  const bool ch_ok = ch0 < p.E;
  const int t_first = chunk * TC + seg_id * L;
  const size_t row0 = (size_t)b * p.T;

  // per-lane constants
  uint32_t cbx[NV], cba[NV], csp[NV];
  if constexpr (KIND == 0) {
    float fbx[V], fba[V], fsp[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { fbx[i] = 0.f; fba[i] = 0.f; fsp[i] = 0.f; }
    if (ch_ok) {
      if constexpr (BF) {
        if (p.bias_x) { uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.bias_x) + ch0);
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) { fbx[2 * i] = bf_lo(w[i]); fbx[2 * i + 1] = bf_hi(w[i]); } }
        if (p.bias_a) { uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.bias_a) + ch0);
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) { fba[2 * i] = bf_lo(w[i]); fba[2 * i + 1] = bf_hi(w[i]); } }
      } else {
        if (p.bias_x) { float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.bias_x) + ch0);
          fbx[0] = v.x; fbx[1] = v.y; fbx[2] = v.z; fbx[3] = v.w; }
        if (p.bias_a) { float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.bias_a) + ch0);
          fba[0] = v.x; fba[1] = v.y; fba[2] = v.z; fba[3] = v.w; }
      }
#pragma unroll
      for (int i = 0; i < V; i += 4) {
        float4 v = *reinterpret_cast<const float4*>(p.neg8sp + ch0 + i);
        fsp[i] = v.x; fsp[i + 1] = v.y; fsp[i + 2] = v.z; fsp[i + 3] = v.w;
      }
    }
    if constexpr (PACKED) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        cbx[i] = pack_bf2(fbx[2 * i], fbx[2 * i + 1]);   // exact: values are bf16
        cba[i] = pack_bf2(fba[2 * i], fba[2 * i + 1]);
        csp[i] = pack_bf2(fsp[2 * i], fsp[2 * i + 1]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        cbx[i] = __float_as_uint(fbx[i]); cba[i] = __float_as_uint(fba[i]);
        csp[i] = __float_as_uint(fsp[i]);
      }
    }
  }

  // ------------------------------------------------------------ pass 1
  uint32_t st_a[L][NV], st_x[L][NV];   // (a_t, x~_t): packed bf16x2 or fp32 bits
  float P[V], H[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { P[i] = 1.0f; H[i] = 0.0f; }

#pragma unroll
  for (int j = 0; j < L; ++j) {
    const int t = t_first + j;
    const bool ok = ch_ok && t < p.T;
    uint4 vx = make_uint4(0, 0, 0, 0), v1 = vx, v2 = vx;
    bool rs = false;
    if (ok) {
      const size_t row = row0 + t;
      vx = ldg_stream(reinterpret_cast<const IO*>(p.x) + row * p.E + ch0);
      if constexpr (KIND == 0) {
        v1 = ldg_stream(reinterpret_cast<const IO*>(p.gemm_x) + row * p.gate_ld + ch0);
        v2 = ldg_stream(reinterpret_cast<const IO*>(p.gemm_a) + row * p.gate_ld + ch0);
        rs = load_seg(p.seg, p.seg_is_i64 != 0, (long long)b * p.seg_bstride + t) == 0;
      } else {
        v1 = ldg_stream(reinterpret_cast<const IO*>(p.a) + row * p.E + ch0);
        rs = p.reset[row] != 0;
      }
    }
    const uint32_t wx[4] = {vx.x, vx.y, vx.z, vx.w};
    const uint32_t w1[4] = {v1.x, v1.y, v1.z, v1.w};
    const uint32_t w2[4] = {v2.x, v2.y, v2.z, v2.w};
    if constexpr (PACKED) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        uint32_t av, nx;
        if constexpr (KIND == 0) {
          gate_pair_emul<FAST>(wx[i], w1[i], w2[i], cbx[i], cba[i], csp[i], rs, av, nx);
        } else {
          av = rs ? 0u : w1[i];   // a * ~reset (:173)
          nx = wx[i];
        }
        if (!ok) { av = kOne2; nx = 0u; }   // identity step beyond T / E
        st_a[j][i] = av; st_x[j][i] = nx;
        const float a0 = bf_lo(av), a1 = bf_hi(av);
        H[2 * i] = fmaf(a0, H[2 * i], bf_lo(nx));
        H[2 * i + 1] = fmaf(a1, H[2 * i + 1], bf_hi(nx));
        P[2 * i] *= a0; P[2 * i + 1] *= a1;
      }
    } else {
      float fx[V], f1[V], f2[V];
      if constexpr (BF) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          fx[2 * i] = bf_lo(wx[i]); fx[2 * i + 1] = bf_hi(wx[i]);
          f1[2 * i] = bf_lo(w1[i]); f1[2 * i + 1] = bf_hi(w1[i]);
          f2[2 * i] = bf_lo(w2[i]); f2[2 * i + 1] = bf_hi(w2[i]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          fx[i] = __uint_as_float(wx[i]); f1[i] = __uint_as_float(w1[i]);
          f2[i] = __uint_as_float(w2[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float av, nx;
        if constexpr (KIND == 0) {
          gate_f32<FAST>(fx[i], f1[i], f2[i], __uint_as_float(cbx[i]), __uint_as_float(cba[i]),
                         __uint_as_float(csp[i]), rs, av, nx);
        } else {
          av = rs ? 0.0f : f1[i];
          nx = fx[i];
        }
        if (!ok) { av = 1.0f; nx = 0.0f; }
        st_a[j][i] = __float_as_uint(av); st_x[j][i] = __float_as_uint(nx);
        H[i] = fmaf(av, H[i], nx);
        P[i] *= av;
      }
    }
  }

  // ------------------------------------------------ warp scan over the 4 segments
#pragma unroll
  for (int d = 8; d <= 16; d <<= 1) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float pp = __shfl_up_sync(0xffffffffu, P[i], d);
      const float hp = __shfl_up_sync(0xffffffffu, H[i], d);
      if (lane >= d) { H[i] = fmaf(P[i], hp, H[i]); P[i] *= pp; }
    }
  }
  float Pex[V], Hex[V];   // transform of the segments before mine
#pragma unroll
  for (int i = 0; i < V; ++i) {
    Pex[i] = __shfl_up_sync(0xffffffffu, P[i], 8);
    Hex[i] = __shfl_up_sync(0xffffffffu, H[i], 8);
    if (seg_id == 0) { Pex[i] = 1.0f; Hex[i] = 0.0f; }
  }
  // lanes 24..31 now hold the whole chunk's (P,H).

  // ------------------------------------------------ carry across chunks
  float c[V];
  const size_t ws_off = (size_t)item * EC + cv * V;
  if (chunk == 0) {
#pragma unroll
    for (int i = 0; i < V; ++i) c[i] = 0.0f;
    if (p.h0 != nullptr && ch_ok) {
#pragma unroll
      for (int i = 0; i < V; i += 4) {
        float4 v = *reinterpret_cast<const float4*>(p.h0 + (size_t)b * p.E + ch0 + i);
        c[i] = v.x; c[i + 1] = v.y; c[i + 2] = v.z; c[i + 3] = v.w;
      }
    }
  } else {
    int j = chunk - 1;
    int f = ld_acquire(p.flags + (size_t)j * p.ncols + col);
    if (f != 2) {
      // predecessor state not final yet: publish my aggregate so successors
      // can fold over it, then look back.
      if (seg_id == kSegs - 1) {
#pragma unroll
        for (int i = 0; i < V; i += 4) {
          stg_cg(p.agg_p + ws_off + i, make_uint4(__float_as_uint(P[i]), __float_as_uint(P[i + 1]),
                                                  __float_as_uint(P[i + 2]), __float_as_uint(P[i + 3])));
          stg_cg(p.agg_h + ws_off + i, make_uint4(__float_as_uint(H[i]), __float_as_uint(H[i + 1]),
                                                  __float_as_uint(H[i + 2]), __float_as_uint(H[i + 3])));
        }
        __threadfence();
      }
      __syncwarp();
      if (lane == 0) st_release(p.flags + item, 1);
      for (;;) {
        f = ld_acquire(p.flags + (size_t)j * p.ncols + col);
        if (f == 2) break;
        if (f == 1) { --j; continue; }   // chunk 0 always ends at 2, so j >= 0
        __nanosleep(64);
      }
    }
    // c = state after chunk j, then fold aggregates j+1 .. chunk-1 in order
    {
      const float* src = p.pref + ((size_t)j * p.ncols + col) * EC + cv * V;
#pragma unroll
      for (int i = 0; i < V; i += 4) {
        uint4 v = ldg_cg(src + i);
        c[i] = __uint_as_float(v.x); c[i + 1] = __uint_as_float(v.y);
        c[i + 2] = __uint_as_float(v.z); c[i + 3] = __uint_as_float(v.w);
      }
    }
    for (int k = j + 1; k < chunk; ++k) {
      const size_t off = ((size_t)k * p.ncols + col) * EC + cv * V;
#pragma unroll
      for (int i = 0; i < V; i += 4) {
        uint4 vp = ldg_cg(p.agg_p + off + i);
        uint4 vh = ldg_cg(p.agg_h + off + i);
        c[i] = fmaf(__uint_as_float(vp.x), c[i], __uint_as_float(vh.x));
        c[i + 1] = fmaf(__uint_as_float(vp.y), c[i + 1], __uint_as_float(vh.y));
        c[i + 2] = fmaf(__uint_as_float(vp.z), c[i + 2], __uint_as_float(vh.z));
        c[i + 3] = fmaf(__uint_as_float(vp.w), c[i + 3], __uint_as_float(vh.w));
      }
    }
  }
  // publish the state after this chunk (not needed after the last one)
  if (chunk + 1 < p.nchunks) {
    if (seg_id == kSegs - 1) {
#pragma unroll
      for (int i = 0; i < V; i += 4) {
        stg_cg(p.pref + ws_off + i,
               make_uint4(__float_as_uint(fmaf(P[i], c[i], H[i])),
                          __float_as_uint(fmaf(P[i + 1], c[i + 1], H[i + 1])),
                          __float_as_uint(fmaf(P[i + 2], c[i + 2], H[i + 2])),
                          __float_as_uint(fmaf(P[i + 3], c[i + 3], H[i + 3]))));
      }
      __threadfence();
    }
    __syncwarp();
    if (lane == 0) st_release(p.flags + item, 2);
  }

  // ------------------------------------------------------------ pass 2 (replay)
  float h[V];
#pragma unroll
  for (int i = 0; i < V; ++i) h[i] = fmaf(Pex[i], c[i], Hex[i]);

#pragma unroll
  for (int j = 0; j < L; ++j) {
    const int t = t_first + j;
    const bool ok = ch_ok && t < p.T;
    uint4 out;
    if constexpr (PACKED) {
      uint32_t o[4];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const uint32_t av = st_a[j][i], nx = st_x[j][i];
        if constexpr (FAST) {
          h[2 * i] = fmaf(bf_lo(av), h[2 * i], bf_lo(nx));
          h[2 * i + 1] = fmaf(bf_hi(av), h[2 * i + 1], bf_hi(nx));
        } else {   // mul then add, as the reference loop (:196)
          h[2 * i] = __fadd_rn(__fmul_rn(bf_lo(av), h[2 * i]), bf_lo(nx));
          h[2 * i + 1] = __fadd_rn(__fmul_rn(bf_hi(av), h[2 * i + 1]), bf_hi(nx));
        }
        o[i] = pack_bf2(h[2 * i], h[2 * i + 1]);
      }
      out = make_uint4(o[0], o[1], o[2], o[3]);
      if (ok) stg_stream(reinterpret_cast<IO*>(p.y) + (row0 + t) * p.E + ch0, out);
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float av = __uint_as_float(st_a[j][i]), nx = __uint_as_float(st_x[j][i]);
        if constexpr (FAST) h[i] = fmaf(av, h[i], nx);
        else h[i] = __fadd_rn(__fmul_rn(av, h[i]), nx);
      }
      if constexpr (BF) {
        out = make_uint4(pack_bf2(h[0], h[1]), pack_bf2(h[2], h[3]),
                         pack_bf2(h[4 % V], h[5 % V]), pack_bf2(h[6 % V], h[7 % V]));
      } else {
        out = make_uint4(__float_as_uint(h[0]), __float_as_uint(h[1]),
                         __float_as_uint(h[2]), __float_as_uint(h[3]));
      }
      if (ok) stg_stream(reinterpret_cast<IO*>(p.y) + (row0 + t) * p.E + ch0, out);
    }
  }
  // hidden state after the last valid step (padding steps are identities)
  if (p.last_h != nullptr && chunk == p.nchunks - 1 && seg_id == kSegs - 1 && ch_ok) {
#pragma unroll
    for (int i = 0; i < V; i += 4)
      *reinterpret_cast<float4*>(p.last_h + (size_t)b * p.E + ch0 + i) =
          make_float4(h[i], h[i + 1], h[i + 2], h[i + 3]);
  }
}

// ---------------------------------------------------------------------------
// -8 * softplus(a_param): once per call, E threads.  Accurate libdevice math.
// emulate != 0 (bf16 reference mode): round softplus to bf16 first (:352).
// ---------------------------------------------------------------------------
__global__ void softplus_param_kernel(const void* a_param, float* out, int E, int is_bf16,
                                      int emulate) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  float ap = is_bf16 ? __uint_as_float((uint32_t)reinterpret_cast<const uint16_t*>(a_param)[e] << 16)
                     : reinterpret_cast<const float*>(a_param)[e];
  float sp = softplus_f(ap);
  if (is_bf16 && emulate) sp = round_bf(sp);
  out[e] = -8.0f * sp;
}

// ---------------------------------------------------------------------------
// Strict sequential kernels: one thread per (b, channel), reference op order,
// no FMA contraction.  Device-side oracle for the chunked kernels.
// ---------------------------------------------------------------------------
template <typename IO>
__device__ __forceinline__ float load_io(const void* p, size_t i) {
  if constexpr (IoVec<IO>::kBf16)
    return __uint_as_float((uint32_t)reinterpret_cast<const uint16_t*>(p)[i] << 16);
  else
    return reinterpret_cast<const float*>(p)[i];
}
template <typename IO>
__device__ __forceinline__ void store_io(void* p, size_t i, float v) {
  if constexpr (IoVec<IO>::kBf16)
    reinterpret_cast<uint16_t*>(p)[i] = (uint16_t)(pack_bf2(v, v) & 0xffffu);
  else
    reinterpret_cast<float*>(p)[i] = v;
}

template <typename IO, int KIND, int ARITH>
__global__ void strict_scan_kernel(const ScanParams p) {
  constexpr bool BF = IoVec<IO>::kBf16;
  constexpr bool EMUL = BF && (ARITH & 1) == 0;
  constexpr bool FAST = (ARITH & 2) != 0;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (e >= p.E) return;
  float h = p.h0 ? p.h0[(size_t)b * p.E + e] : 0.0f;
  float bx = 0.f, ba = 0.f, sp8 = 0.f;
  if constexpr (KIND == 0) {
    if (p.bias_x) bx = load_io<IO>(p.bias_x, e);
    if (p.bias_a) ba = load_io<IO>(p.bias_a, e);
    sp8 = p.neg8sp[e];
  }
  for (int t = 0; t < p.T; ++t) {
    const size_t row = (size_t)b * p.T + t;
    const float xv = load_io<IO>(p.x, row * p.E + e);
    float av, nx;
    if constexpr (KIND == 0) {
      const bool rs = load_seg(p.seg, p.seg_is_i64 != 0, (long long)b * p.seg_bstride + t) == 0;
      const float g1 = load_io<IO>(p.gemm_x, row * p.gate_ld + e);
      const float g2 = load_io<IO>(p.gemm_a, row * p.gate_ld + e);
      if constexpr (EMUL) {
        // scalar spelling of gate_pair_emul: round after every eager op
        const float px = round_bf(g1 + bx), pa = round_bf(g2 + ba);
        float gx, ga;
        sigmoid2<FAST>(px, pa, gx, ga);
        gx = round_bf(gx); ga = round_bf(ga);
        const float la = round_bf(ga * sp8);
        const float ev = exp_f<FAST>(la);
        float q;
        if constexpr (FAST) q = ev * ev; else q = expf(2.0f * la);
        av = round_bf(ev);
        float mu = round_bf(sqrt_f<FAST>(round_bf(1.0f - round_bf(q))));
        if (rs) { mu = 1.0f; av = 0.0f; }
        nx = round_bf(round_bf(xv * gx) * mu);
      } else {
        gate_f32<FAST>(xv, g1, g2, bx, ba, sp8, rs, av, nx);
      }
    } else {
      av = p.reset[row] ? 0.0f : load_io<IO>(p.a, row * p.E + e);
      nx = xv;
    }
    if (p.T == 1 && p.h0 == nullptr) h = nx;          // :177-178
    else h = __fadd_rn(__fmul_rn(av, h), nx);          // :196 / :181
    store_io<IO>(p.y, row * p.E + e, h);
  }
  if (p.last_h) p.last_h[(size_t)b * p.E + e] = h;
}

}  // namespace cg
