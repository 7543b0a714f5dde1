// Training-path kernels of the RG-LRU gate math.
//
// Forward  (reference recurrentgemma/torch/layers.py:348-365): from x and the two
//   gate pre-activations (BlockDiagonalLinear outputs, bias included) produce the
//   scan's inputs  a = exp(-8 sigmoid(pre_a) softplus(a_param))  and
//   x~ = x sigmoid(pre_x) sqrt(1 - a^2)  (multiplier 1 at a document start).
//   The eager bf16 rounding points are reproduced (gate_pair_emul, exact math).
// Backward (what autograd derives from those lines, with the clipped square-root
//   gradient of SqrtBoundDerivative, :224-238): from grad(x~) and grad(a) -- the
//   outputs of cg_rnn_scan_bwd -- to grad(x), grad(pre_x), grad(pre_a) and
//   grad(a_param).  One pass; the forward's saved tensors are recomputed from the
//   pre-activations WITH their eager rounding, the chain rule runs in fp32;
//   grad(a_param) is reduced in a fixed order (registers -> shared memory -> one
//   partial row per CTA -> column_reduce_kernel).
#pragma once

#include "cg_common.cuh"
#include "cg_scan.cuh"

namespace cg {

constexpr float kMaxSqrtGradient = 1000.0f;   // layers.py:32
constexpr int kTrainRows = 64;                // time steps per CTA of the backward kernel

struct GateParams {
  const void* x;        // [N,E]  (N = B*T rows)
  const void* pre_x;    // [N,E]  input-gate pre-activation, bias included
  const void* pre_a;    // [N,E]
  const void* a_param;  // [E]
  const unsigned char* reset;   // [N]
  // forward outputs
  void* a;              // [N,E]
  void* nx;             // [N,E]
  // backward inputs / outputs
  const void* d_nx;     // [N,E]
  const void* d_a;      // [N,E]
  void* dx;             // [N,E]
  void* d_pre_x;        // [N,E]
  void* d_pre_a;        // [N,E]
  float* partial;       // [ceil(N / kTrainRows)][E] partial sums of grad(a_param)
  long long N;
  int E;
};

template <typename IO>
__device__ __forceinline__ uint4 narrow_vec(const float (&f)[IoVec<IO>::V]) {
  if constexpr (IoVec<IO>::kBf16)
    return make_uint4(pack_bf2(f[0], f[1]), pack_bf2(f[2], f[3]), pack_bf2(f[4 % IoVec<IO>::V], f[5 % IoVec<IO>::V]),
                      pack_bf2(f[6 % IoVec<IO>::V], f[7 % IoVec<IO>::V]));
  else
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
}
template <typename IO>
__device__ __forceinline__ void widen16(const uint4& v, float (&f)[IoVec<IO>::V]) {
  if constexpr (IoVec<IO>::kBf16) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[2 * i] = bf_lo(w[i]); f[2 * i + 1] = bf_hi(w[i]); }
  } else {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
    f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
  }
}

// ---- forward: thread = one 16-byte channel vector of one row
template <typename IO>
__global__ void __launch_bounds__(256)
rglru_gates_fwd_kernel(const GateParams p) {
  constexpr int V = IoVec<IO>::V;
  constexpr bool BF = IoVec<IO>::kBf16;
  const int vecs = p.E / V;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.N * vecs) return;
  const long long row = idx / vecs;
  const int ch0 = (int)(idx - row * vecs) * V;
  const size_t off = (size_t)row * p.E + ch0;
  const uint4 vx = ldg_stream(reinterpret_cast<const IO*>(p.x) + off);
  const uint4 v1 = ldg_stream(reinterpret_cast<const IO*>(p.pre_x) + off);
  const uint4 v2 = ldg_stream(reinterpret_cast<const IO*>(p.pre_a) + off);
  const bool rs = p.reset[row] != 0;
  uint4 oa, on;
  if constexpr (BF) {
    const uint32_t wx[4] = {vx.x, vx.y, vx.z, vx.w}, w1[4] = {v1.x, v1.y, v1.z, v1.w}, w2[4] = {v2.x, v2.y, v2.z, v2.w};
    uint32_t ra[4], rn[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      // -8 * r(softplus(a_param)): softplus rounded to bf16 first (:352), x8 exact
      const uint32_t apw = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint16_t*>(p.a_param) + ch0)[i];
      const float s0 = -8.0f * round_bf(softplus_f(bf_lo(apw))), s1 = -8.0f * round_bf(softplus_f(bf_hi(apw)));
      const uint32_t sp8 = pack_bf2(s0, s1);
      uint32_t av, nv;
      gate_pair_emul<false, false>(wx[i], w1[i], w2[i], 0u, 0u, sp8, av, nv);
      if (rs) {               // multiplier 1 at a document start (:364); `a` itself stays (rnn_scan zeroes it, :173)
        uint32_t az, nr;
        gate_pair_emul<false, true>(wx[i], w1[i], w2[i], 0u, 0u, sp8, az, nr);
        nv = nr;
      }
      ra[i] = av; rn[i] = nv;
    }
    oa = make_uint4(ra[0], ra[1], ra[2], ra[3]);
    on = make_uint4(rn[0], rn[1], rn[2], rn[3]);
  } else {
    float fx[V], f1[V], f2[V], fa[V], fn[V];
    widen16<IO>(vx, fx); widen16<IO>(v1, f1); widen16<IO>(v2, f2);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float sp8 = -8.0f * softplus_f(reinterpret_cast<const float*>(p.a_param)[ch0 + i]);
      float av, nv;
      gate_f32<false, false>(fx[i], f1[i], f2[i], 0.0f, 0.0f, sp8, av, nv);
      if (rs) { float az; gate_f32<false, true>(fx[i], f1[i], f2[i], 0.0f, 0.0f, sp8, az, nv); }
      fa[i] = av; fn[i] = nv;
    }
    oa = narrow_vec<IO>(fa);
    on = narrow_vec<IO>(fn);
  }
  stg_stream(reinterpret_cast<IO*>(p.a) + off, oa);
  stg_stream(reinterpret_cast<IO*>(p.nx) + off, on);
}

// ---- backward: CTA = 128 threads = 8 lanes (one 128-byte row segment... of
// 8 x V channels) x 16 row slots; each thread walks kTrainRows / 16 rows.
template <typename IO>
__global__ void __launch_bounds__(128)
rglru_gates_bwd_kernel(const GateParams p) {
  constexpr int V = IoVec<IO>::V;
  constexpr int EC = kCvl * V;
  __shared__ float red[16][kCvl][V + 1];
  const int cv = threadIdx.x & 7;
  const int slot = threadIdx.x >> 3;
  const int ch0 = blockIdx.x * EC + cv * V;
  const long long row0 = (long long)blockIdx.y * kTrainRows;
  float dap[V];
#pragma unroll
  for (int i = 0; i < V; ++i) dap[i] = 0.0f;
  if (ch0 < p.E) {
    float sp[V], sg[V];                         // softplus(a_param), its derivative sigmoid(a_param)
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float ap = load_io<IO>(p.a_param, ch0 + i);
      sp[i] = softplus_f(ap);
      if constexpr (IoVec<IO>::kBf16) sp[i] = round_bf(sp[i]);
      sg[i] = 1.0f / (1.0f + expf(-ap));
    }
#pragma unroll 1
    for (int r = slot; r < kTrainRows; r += 16) {
      const long long row = row0 + r;
      if (row >= p.N) break;
      const size_t off = (size_t)row * p.E + ch0;
      const uint4 vx = ldg_stream(reinterpret_cast<const IO*>(p.x) + off);
      const uint4 v1 = ldg_stream(reinterpret_cast<const IO*>(p.pre_x) + off);
      const uint4 v2 = ldg_stream(reinterpret_cast<const IO*>(p.pre_a) + off);
      const uint4 vn = ldg_stream(reinterpret_cast<const IO*>(p.d_nx) + off);
      const uint4 va = ldg_stream(reinterpret_cast<const IO*>(p.d_a) + off);
      const bool rs = p.reset[row] != 0;
      float x[V], px[V], pa[V], dnx[V], da[V], dx[V], dpx[V], dpa[V];
      widen16<IO>(vx, x); widen16<IO>(v1, px); widen16<IO>(v2, pa);
      widen16<IO>(vn, dnx); widen16<IO>(va, da);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        // the tensors autograd saved in the forward, with their eager rounding
        // (bf16): where a^2 rounds to 1 the saved 1 - a^2 is exactly 0 and the
        // clipped square-root gradient saturates -- that must be reproduced
        auto rb = [](float v) { if constexpr (IoVec<IO>::kBf16) return round_bf(v); else return v; };
        const float gx = rb(1.0f / (1.0f + expf(-px[i])));
        const float ga = rb(1.0f / (1.0f + expf(-pa[i])));
        const float log_a = rb(-8.0f * ga * sp[i]);
        const float a = rb(expf(log_a));
        const float a2 = rb(expf(2.0f * log_a));
        const float om = rb(1.0f - a2);
        const float mult = rs ? 1.0f : rb(sqrtf(om));
        const float gated = rb(x[i] * gx);
        const float d_gated = dnx[i] * mult;
        const float d_mult = rs ? 0.0f : dnx[i] * gated;
        dx[i] = d_gated * gx;
        const float d_gx = d_gated * x[i];
        // clipped gradient of the square root (:236-238)
        const float d_om = d_mult / sqrtf(fmaxf(4.0f * om, 1.0f / (kMaxSqrtGradient * kMaxSqrtGradient)));
        const float d_log_a = a * da[i] - 2.0f * a2 * d_om;
        const float d_ga = -8.0f * sp[i] * d_log_a;
        dap[i] += -8.0f * ga * d_log_a * sg[i];
        dpa[i] = d_ga * ga * (1.0f - ga);
        dpx[i] = d_gx * gx * (1.0f - gx);
      }
      stg_stream(reinterpret_cast<IO*>(p.dx) + off, narrow_vec<IO>(dx));
      stg_stream(reinterpret_cast<IO*>(p.d_pre_x) + off, narrow_vec<IO>(dpx));
      stg_stream(reinterpret_cast<IO*>(p.d_pre_a) + off, narrow_vec<IO>(dpa));
    }
  }
#pragma unroll
  for (int i = 0; i < V; ++i) red[slot][cv][i] = dap[i];
  __syncthreads();
  for (int idx = threadIdx.x; idx < kCvl * V; idx += blockDim.x) {
    const int c = idx / V, i = idx - c * V;
    const int ch = blockIdx.x * EC + c * V + i;
    if (ch >= p.E) continue;
    float sum = 0.0f;
#pragma unroll
    for (int sl = 0; sl < 16; ++sl) sum += red[sl][c][i];
    p.partial[(size_t)blockIdx.y * p.E + ch] = sum;
  }
}

// out[e] = sum over parts of partial[part][e], in part order (fixed: reproducible)
template <typename IO>
__global__ void column_reduce_kernel(const float* __restrict__ partial, int nparts, int E, void* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
  int q = 0;
  for (; q + 3 < nparts; q += 4) {
    s0 += partial[(size_t)q * E + e];
    s1 += partial[(size_t)(q + 1) * E + e];
    s2 += partial[(size_t)(q + 2) * E + e];
    s3 += partial[(size_t)(q + 3) * E + e];
  }
  for (; q < nparts; ++q) s0 += partial[(size_t)q * E + e];
  store_io<IO>(out, e, (s0 + s1) + (s2 + s3));
}

}  // namespace cg
