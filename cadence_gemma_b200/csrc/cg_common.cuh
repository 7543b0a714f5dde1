// Device helpers shared by the sm_100a kernels of the recurrent hot path.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace cg {

constexpr uint32_t kOne2 = 0x3f803f80u;  // bf16x2 {1.0, 1.0}
constexpr float kLog2e = 1.4426950408889634f;

// ---------------------------------------------------------------- bf16 <-> f32
__device__ __forceinline__ float bf_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
// {lo, hi} fp32 -> packed bf16x2, round-to-nearest-even (what ATen's cast does).
__device__ __forceinline__ uint32_t pack_bf2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// fp32 -> nearest bf16 -> fp32 (one eager rounding point of the reference).
__device__ __forceinline__ float round_bf(float v) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(v), "f"(v));
  return __uint_as_float(d & 0xffff0000u);
}

// Packed bf16 arithmetic with ONE rounding per op and no contraction.  The
// reference evaluates bf16 elementwise ops as fp32 op + round-to-bf16; since
// fp32 carries >= 2*8+2 significand bits that double rounding is innocuous
// for +,-,*,sqrt, i.e. it equals the correctly rounded bf16 op computed here.
__device__ __forceinline__ uint32_t bf2_mul(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t bf2_add(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t bf2_sub(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("sub.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

// ------------------------------------------------------------- transcendentals
// CG_ABLATE_MUFU (timing experiments only, results are WRONG): 1 = no MUFU.SQRT,
// 2 = no MUFU.EX2, 4 = no MUFU.RCP
#ifndef CG_ABLATE_MUFU
#define CG_ABLATE_MUFU 0
#endif
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  if (CG_ABLATE_MUFU & 2) return x * 0.125f + 0.5f;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  if (CG_ABLATE_MUFU & 4) return 1.0f - x * 0.01f;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  if (CG_ABLATE_MUFU & 1) return x * 0.75f;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool FAST>
__device__ __forceinline__ float exp_f(float x) {
  if constexpr (FAST) return ex2_approx(x * kLog2e);
  else return expf(x);
}
template <bool FAST>
__device__ __forceinline__ float sqrt_f(float x) {
  if constexpr (FAST) return sqrt_approx(x);
  else return sqrtf(x);
}
// sigmoid of two values.  Exact: the ATen formula 1/(1+exp(-v)) with IEEE
// division.  Fast: ex2.approx and ONE shared reciprocal,
// r = 1/((1+e0)(1+e1)), s0 = r(1+e1), s1 = r(1+e0): the MUFU unit (16 lanes /
// clk / SM) is the busiest pipe of the fused kernel, so trading a MUFU.RCP for
// three FMULs pays.  Callers clamp the arguments at -43 (packed, one HMNMX2
// per pair) so that (1+e0)(1+e1) <= 2^124 stays finite.
template <bool FAST>
__device__ __forceinline__ void sigmoid2(float v0, float v1, float& s0, float& s1) {
  if constexpr (FAST) {
    const float d0 = 1.0f + ex2_approx(v0 * -kLog2e);
    const float d1 = 1.0f + ex2_approx(v1 * -kLog2e);
    const float r = rcp_approx(d0 * d1);
    s0 = r * d1;
    s1 = r * d0;
  } else {
    s0 = 1.0f / (1.0f + expf(-v0));
    s1 = 1.0f / (1.0f + expf(-v1));
  }
}
// max(v, -43) on a bf16x2 pair (one instruction for two values).
__device__ __forceinline__ uint32_t bf2_clamp_lo(uint32_t v) {
  uint32_t d;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(v), "r"(0xc22cc22cu));   // -43.0 | -43.0
  return d;
}
// F.softplus (beta 1, threshold 20) -- parameters only, always accurate.
__device__ __forceinline__ float softplus_f(float v) {
  return v > 20.0f ? v : log1pf(expf(v));
}

// ------------------------------------------------------------ global memory I/O
// Streaming 128-bit accesses: inputs are read once (no L1 allocation), outputs
// are written once (evict-first).
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(void* p, uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// L2-coherent accesses for the cross-warp carry exchange.
__device__ __forceinline__ uint4 ldg_cg(const void* p) {
  uint4 r;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void stg_cg(void* p, uint4 v) {
  asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// 64-bit exchange words {fp32 value : high 32, tag : low 32}.  An aligned
// 8-byte access is single-copy atomic, so value and tag always travel together.
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long pack_tagged(float v, unsigned tag) {
  return ((unsigned long long)__float_as_uint(v) << 32) | tag;
}
__device__ __forceinline__ float tagged_value(unsigned long long w) {
  return __uint_as_float((unsigned)(w >> 32));
}

// cp.async (LDGSTS): 16-byte global -> shared copy that bypasses L1 and holds
// no registers while in flight.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(gmem_src) : "memory");
}
// Prefetch the 128-byte line containing `gmem_src` into L2 (per-thread address;
// the bulk form needs warp-uniform operands and would serialise over lanes).
__device__ __forceinline__ void prefetch_l2_line(const void* gmem_src) {
  asm volatile("prefetch.global.L2 [%0];" :: "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory");
}
// wait until at most `pending` of this thread's groups are outstanding
// (`pending` is a compile-time value after unrolling; MAXN bounds the switch).
template <int MAXN>
__device__ __forceinline__ void cp_async_wait_dyn(int pending) {
  if constexpr (MAXN > 0) {
    if (pending >= MAXN - 1) { cp_async_wait<MAXN - 1>(); return; }
    cp_async_wait_dyn<MAXN - 1>(pending);
  } else {
    cp_async_wait<0>();
  }
}

// segment_pos[idx] == 0 (a document start), int32 or int64 positions
__device__ __forceinline__ bool seg_is_zero(const void* seg, bool is_i64, long long idx) {
  if (is_i64) {
    const uint2 v = reinterpret_cast<const uint2*>(seg)[idx];
    return (v.x | v.y) == 0u;
  }
  return reinterpret_cast<const int*>(seg)[idx] == 0;
}
__device__ __forceinline__ long long load_seg(const void* seg, bool is_i64, long long idx) {
  return is_i64 ? reinterpret_cast<const long long*>(seg)[idx]
                : static_cast<long long>(reinterpret_cast<const int*>(seg)[idx]);
}

}  // namespace cg
