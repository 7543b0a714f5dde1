// Chunked two-level linear-recurrence scan for sm_100a, fused with the RG-LRU
// gate math.  Replaces reference recurrentgemma/torch/layers.py:345-375
// (RGLRU.forward after the gate GEMMs) and :146-199 (rnn_scan).
//
// Decomposition:
//   CTA    = persistent, NW warps.  Work item = NW consecutive time chunks (a
//            "super-chunk") of one channel tile (batch row b, EC = 8*V channels
//            = one 128-byte row), taken from an atomic ticket in time-major
//            order (adjacent tickets = adjacent column tiles of the same rows).
//   warp   = one chunk of TC = 4*L steps.  lane = (seg = lane / 8, cv = lane % 8):
//            8 lanes x V channels (16-byte vectors) cover one contiguous 128-byte
//            row, the 4 lane groups take 4 consecutive time segments of L steps
//            => every warp-wide access moves 4 full 128 B lines.
//   stage  : each lane copies its own L steps of x / gate pre-activations
//            global -> shared with cp.async (LDGSTS, 16 B, L1 bypass) into one
//            of two buffers: the next item streams in while the current one is
//            computed, and the bytes in flight live in shared memory, not
//            registers.
//   pass 1 : lanes evaluate the gates in registers (sigmoid, softplus-exp,
//            sqrt(1-a^2), every bf16 rounding point of the reference), write
//            (x~_t, a_t) back over the staged inputs and accumulate their
//            segment's transform h -> P*h + H in fp32.
//   carry  : the last warp scans the item's 4*NW segment transforms in shared
//            memory; items of one column are chained through global memory with
//            a decoupled look-back (flag 1 = aggregate published, 2 = state
//            published).  Carries are always folded left-to-right over
//            aggregates, so results do not depend on timing (bit-reproducible).
//   pass 2 : replay h = a_t*h + x~_t from the true carry-in, store y (bf16/fp32).
// A claimed ticket belongs to a resident CTA that works its tickets in
// increasing order and dependencies only point to lower tickets, so the
// look-back cannot deadlock.
#pragma once

#include "cg_common.cuh"

namespace cg {

constexpr int kP1Unroll = 2;   // pass-1 steps interleaved per lane (ILP vs instruction-cache footprint)

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
  const uint16_t* neg8sp_bf;   // [E] the same as bf16 (exact in reference mode)
  const unsigned* reset_bits;  // [rows][ceil(T/32)] bit t%32 of word t/32 = reset at t
  long long bits_bstride;      // words per batch row (0 = one row broadcast)
  const void* seg;      // segment positions
  long long seg_bstride;
  int seg_is_i64;
  long long gate_ld;
  int gate_bw;          // > 0: gate GEMM outputs are interleaved per head, see gate_off()
  // plain scan inputs (KIND 1): x above, plus
  const void* a;                 // [B,T,E]
  const unsigned char* reset;    // [B,T]
  // backward of the plain scan (KIND 2): x = grad of y, a / reset as forward, plus
  const void* hprev;             // [B,T,E] forward output y (h_t in the I/O dtype); h0 above = forward h0
  const float* g_last;           // [B,E] grad of last_h or null
  void* da;                      // [B,T,E] grad of a; `y` above receives grad of x; last_h receives grad of h0
  // state
  const float* h0;   // [B,E] or null
  void* y;           // [B,T,E]
  float* last_h;     // [B,E] or null
  // scratch (see cadence_b200.cu: carve()).  The three exchange arrays hold
  // 64-bit words {fp32 value : high, launch epoch : low}; a word is valid for
  // this launch iff its low half equals *epoch, so one relaxed 8-byte access
  // carries data and "ready" flag together (no fence, no separate flag load).
  int* counter;                 // ticket, reset by the prologue kernel
  const unsigned* epoch;        // bumped by the prologue kernel
  unsigned long long* agg_p;    // [nitems][EC] item aggregate  h -> P*h + H
  unsigned long long* agg_h;
  unsigned long long* pref;     // [nitems][EC] state leaving the item
  int B, T, E;
  int ncols;         // B * ceil(E / EC)
  int ctiles;        // ceil(E / EC)
  int nchunks;
  int nitems;
};

struct TrueTag { static constexpr bool value = true; };
struct FalseTag { static constexpr bool value = false; };

// Column of channel `ch` inside a gate-GEMM output row.  gate_bw == 0: plain
// [.., E] rows.  gate_bw == bw > 0: one fused GEMM wrote [.., H, 2*bw] rows
// (input-gate block then a-gate block per head); gemm_a points bw further.
__device__ __forceinline__ int gate_off(int ch, int gate_bw) {
  return gate_bw > 0 ? ch + (ch / gate_bw) * gate_bw : ch;
}

template <typename IO> struct IoVec;
template <> struct IoVec<uint16_t> { static constexpr int V = 8; static constexpr bool kBf16 = true; };
template <> struct IoVec<float> { static constexpr int V = 4; static constexpr bool kBf16 = false; };

// ---- RG-LRU gates on one bf16x2 pair, every eager rounding point reproduced.
// layers.py:348-365 + :173; see SURVEY.md section 7 "bf16 rounding points".
template <bool FAST, bool RESET>
__device__ __forceinline__ void gate_pair_emul(uint32_t xc, uint32_t gxr, uint32_t gar,
                                               uint32_t bx, uint32_t ba, uint32_t sp8,
                                               uint32_t& a_out, uint32_t& nx_out) {
  const uint32_t px = bf2_add(gxr, bx);          // r(gemm + b)            :139
  const uint32_t pa = bf2_add(gar, ba);
  float sx0, sx1, sa0, sa1;
  if constexpr (FAST) {                          // pair each x gate with an a gate
    const uint32_t cx = bf2_clamp_lo(px), ca = bf2_clamp_lo(pa);
    sigmoid2<true>(bf_lo(cx), bf_lo(ca), sx0, sa0);
    sigmoid2<true>(bf_hi(cx), bf_hi(ca), sx1, sa1);
  } else {
    sigmoid2<false>(bf_lo(px), bf_hi(px), sx0, sx1);
    sigmoid2<false>(bf_lo(pa), bf_hi(pa), sa0, sa1);
  }
  const uint32_t gx = pack_bf2(sx0, sx1);        // r(sigmoid)             :348
  if constexpr (RESET) {
    // a document start: a = 0 (:173), multiplier = 1 (:364): x~ = r(x*gx)
    nx_out = bf2_mul(xc, gx);
    a_out = 0u;
    return;
  }
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
  const uint32_t av = pack_bf2(e0, e1);
  const uint32_t om = bf2_sub(kOne2, pack_bf2(q0, q1));   // r(1 - a^2)    :361
  const uint32_t mu = pack_bf2(sqrt_f<FAST>(bf_lo(om)), sqrt_f<FAST>(bf_hi(om)));
  nx_out = bf2_mul(bf2_mul(xc, gx), mu);         // r(r(x*gx)*mu)    :357, :365
  a_out = av;
}

// ---- same, fp32 kept in registers (bf16 or fp32 inputs already widened).
// For fp32 tensors this IS the reference op order (each op rounds to fp32).
template <bool FAST, bool RESET>
__device__ __forceinline__ void gate_f32(float xc, float gxr, float gar, float bx, float ba,
                                         float sp8, float& a_out, float& nx_out) {
  float px = gxr + bx, pa = gar + ba;
  if constexpr (FAST) { px = fmaxf(px, -43.0f); pa = fmaxf(pa, -43.0f); }
  float gx, ga;
  sigmoid2<FAST>(px, pa, gx, ga);
  if constexpr (RESET) {   // a document start: a = 0 (:173), multiplier = 1 (:364)
    nx_out = __fmul_rn(xc, gx);
    a_out = 0.0f;
    return;
  }
  const float la = ga * sp8;
  const float e = exp_f<FAST>(la);
  float q;
  if constexpr (FAST) q = e * e; else q = expf(2.0f * la);
  const float mu = sqrt_f<FAST>(1.0f - q);
  nx_out = __fmul_rn(__fmul_rn(xc, gx), mu);
  a_out = e;
}

// KIND: 0 = RG-LRU (gates fused), 1 = plain rnn_scan(x, a, reset, h0),
// 2 = backward of the plain scan: the same recurrence run over reversed time,
//   dh_t = a_{t+1} * dh_{t+1} + gy_t,   dx_t = dh_t,   da_t = dh_t * h_{t-1}
// (autograd of the reference loop, layers.py:187-199: mul then add in fp32,
// casts at the ends).  Logical step tau of the scan is physical row T-1-tau.
// ARITH: CG_ARITH_* bits 0 (fp32 in registers) and 1 (fast math).
template <typename IO, int KIND, int ARITH>
struct ScanTraits {
  static constexpr int V = IoVec<IO>::V;
  static constexpr bool BF = IoVec<IO>::kBf16;
  static constexpr bool FAST = (ARITH & 2) != 0;
  // (a, x~) are exactly bf16 -> kept packed between the two passes
  static constexpr bool PACKED = BF && (KIND >= 1 || (ARITH & 1) == 0);
  static constexpr int EC = kCvl * V;          // channels per tile (128 B rows)
  static constexpr int CPL = EC / 32;          // channels per lane in the carry warp
  // 16-byte staging slots per (step, lane): inputs x, g1, g2 (or x, a); the
  // (x~, a) state overwrites them in place (fp32 state of 8 channels needs 4)
  static constexpr int NT = KIND == 1 ? 2 : ((BF && !PACKED) ? 4 : 3);   // KIND 2: gy, a, h_prev
  // With 3 slots the third one (gemm_a) is dead after pass 1: the segment
  // transforms live there instead of in a separate shared array.
  static constexpr bool ALIAS = NT == 3 && KIND == 0;
};

// STAGES staging buffers per CTA: with 2, item i+1 streams in while item i is
// computed; with 1 the latency is covered by the other resident CTAs instead.
template <typename IO, int KIND, int ARITH, int L, int NW, int STAGES>
constexpr size_t scan_smem_bytes() {
  using Tr = ScanTraits<IO, KIND, ARITH>;
  return (size_t)STAGES * NW * L * Tr::NT * 512                 // staged inputs / (x~, a) state
         + (Tr::ALIAS ? 0 : (size_t)2 * NW * kSegs * Tr::EC * sizeof(float))   // segment transforms
         + (size_t)2 * (2 * NW + 1) * Tr::EC * sizeof(float)    // warp transforms + carry, 2 parities
         + 128 + 16;                                            // 4 claimed work items, flags
}

// Persistent CTA of NW warps.  Work item = NW consecutive chunks (a super-chunk
// of NW*4*L steps) of one 128-byte-wide channel column; items come from an
// atomic ticket in time-major order.  Two staging buffers: the inputs of the
// next item stream in (cp.async) while the current one is computed.
template <typename IO, int KIND, int ARITH, int L, int NW, int STAGES, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB)
scan_kernel(const ScanParams p) {
  using Tr = ScanTraits<IO, KIND, ARITH>;
  constexpr int V = Tr::V;
  constexpr bool BF = Tr::BF;
  constexpr bool FAST = Tr::FAST;
  constexpr bool PACKED = Tr::PACKED;
  constexpr int EC = Tr::EC;
  constexpr int CPL = Tr::CPL;
  constexpr int NT = Tr::NT;
  constexpr int TC = kSegs * L;
  constexpr int NV = PACKED ? V / 2 : V;
  constexpr int NSEG = NW * kSegs;           // time segments per item
  constexpr int STAGE_U4 = NW * L * NT * 32; // uint4 per staging buffer

  // launched as a programmatic dependent of the prologue kernel (ticket reset,
  // epoch, reset bitmask): block scheduling overlaps the prologue, everything
  // below reads its outputs
  asm volatile("griddepcontrol.wait;" ::: "memory");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint4* stage = reinterpret_cast<uint4*>(smem_raw);
  constexpr bool ALIAS = Tr::ALIAS;
  static_assert(!ALIAS || L >= V / 2, "aliased segment transforms need V/2 steps");
  float* s_p = reinterpret_cast<float*>(smem_raw + (size_t)STAGES * STAGE_U4 * 16);
  float* s_h = s_p + (ALIAS ? 0 : NSEG * EC);
  float* s_wp = s_h + (ALIAS ? 0 : NSEG * EC);    // [2][NW][EC] warp transforms (P)
  float* s_wh = s_wp + 2 * NW * EC;               // [2][NW][EC] warp transforms (H)
  float* s_c0 = s_wh + 2 * NW * EC;               // [2][EC] state entering the item
  int* s_tk = reinterpret_cast<int*>(s_c0 + 2 * EC);   // 4 slots x 8 ints of claimed work items
  int* s_flag = s_tk + 32;                        // [2] carry-in known before the barrier?

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int seg_id = lane >> 3;
  const int cv = lane & 7;
  const int my_seg = warp * kSegs + seg_id;  // segment index inside the item

  // Work-item coordinates, decoded once per ticket by thread 0.
  struct Coord { int item, sc, col, b, e0; };
  auto claim = [&](int slot) {               // thread 0 only
    const int item = atomicAdd(p.counter, 1);
    int sc = 0, col = 0, b = 0, e0 = 0;
    if (item < p.nitems) {
      sc = item / p.ncols;                   // time-major ticket order
      col = item - sc * p.ncols;
      b = col / p.ctiles;
      e0 = (col - b * p.ctiles) * EC;
    }
    int* d = s_tk + slot * 8;
    d[0] = item; d[1] = sc; d[2] = col; d[3] = b; d[4] = e0;
  };
  auto fetch = [&](int slot) {
    const int* d = s_tk + slot * 8;
    Coord c; c.item = d[0]; c.sc = d[1]; c.col = d[2]; c.b = d[3]; c.e0 = d[4];
    return c;
  };
  // number of this lane's L steps that exist (0 beyond T or E)
  auto valid_steps = [&](const Coord& c) {
    if (c.item >= p.nitems || c.e0 + cv * V >= p.E) return 0;
    const int left = p.T - ((c.sc * NW + warp) * TC + seg_id * L);
    return left < 0 ? 0 : (left > L ? L : left);
  };
  // global -> shared copies of this lane's own steps (LDGSTS, 16 B each).
  // Every lane later consumes exactly the bytes it copied itself: no barrier,
  // only cp.async.wait_group.  One commit group per item.
  auto issue_loads = [&](const Coord& c, int buf) {
    const int nv = valid_steps(c);
    if constexpr (KIND == 2) {
      if (nv > 0) {
        // reversed time: logical step tau reads gy[t], a[t+1] (1 at tau == 0: the
        // incoming grad of last_h enters with coefficient one) and h[t-1], t = T-1-tau
        const int ch0 = c.e0 + cv * V;
        const int tau0 = (c.sc * NW + warp) * TC + seg_id * L;
        const size_t rowb = (size_t)c.b * p.T;
        uint4* dst = stage + (size_t)buf * STAGE_U4 + (size_t)warp * L * NT * 32 + lane;
        const uint32_t one = BF ? kOne2 : 0x3f800000u;
        for (int j = 0; j < nv; ++j) {
          const int t = p.T - 1 - (tau0 + j);
          cp_async16(dst, reinterpret_cast<const IO*>(p.x) + (rowb + t) * p.E + ch0);
          if (t + 1 < p.T) cp_async16(dst + 32, reinterpret_cast<const IO*>(p.a) + (rowb + t + 1) * p.E + ch0);
          else dst[32] = make_uint4(one, one, one, one);
          if (t > 0) cp_async16(dst + 64, reinterpret_cast<const IO*>(p.hprev) + (rowb + t - 1) * p.E + ch0);
          dst += NT * 32;
        }
      }
      cp_async_commit();
      return;
    }
    if (nv > 0) {
      const int ch0 = c.e0 + cv * V;
      const size_t row = (size_t)c.b * p.T + (c.sc * NW + warp) * TC + seg_id * L;
      uint4* dst = stage + (size_t)buf * STAGE_U4 + (size_t)warp * L * NT * 32 + lane;
      const IO* px = reinterpret_cast<const IO*>(p.x) + row * p.E + ch0;
      const int g0 = KIND == 0 ? gate_off(ch0, p.gate_bw) : ch0;
      const IO* p1 = KIND == 0 ? reinterpret_cast<const IO*>(p.gemm_x) + row * p.gate_ld + g0
                               : reinterpret_cast<const IO*>(p.a) + row * p.E + ch0;
      const IO* p2 = reinterpret_cast<const IO*>(p.gemm_a) + row * p.gate_ld + g0;
      const size_t ld1 = KIND == 0 ? (size_t)p.gate_ld : (size_t)p.E;
#pragma unroll 2
      for (int j = 0; j < nv; ++j) {
        cp_async16(dst, px);
        cp_async16(dst + 32, p1);
        if constexpr (KIND == 0) cp_async16(dst + 64, p2);
        dst += NT * 32; px += p.E; p1 += ld1; p2 += ld1;
      }
    }
    cp_async_commit();
  };

  // L2 prefetch of an item's input rows (one 128-byte row per lane and tensor),
  // issued as soon as the item is known: the later cp.async then hit L2.
  auto prefetch_l2 = [&](const Coord& c) {
    if (KIND == 2 || c.item >= p.nitems || lane >= TC) return;
    const int t = (c.sc * NW + warp) * TC + lane;
    if (t >= p.T) return;
    const size_t row = (size_t)c.b * p.T + t;
    prefetch_l2_line(reinterpret_cast<const IO*>(p.x) + row * p.E + c.e0);
    if constexpr (KIND == 0) {
      const int g0 = gate_off(c.e0, p.gate_bw);
      prefetch_l2_line(reinterpret_cast<const IO*>(p.gemm_x) + row * p.gate_ld + g0);
      prefetch_l2_line(reinterpret_cast<const IO*>(p.gemm_a) + row * p.gate_ld + g0);
    } else {
      prefetch_l2_line(reinterpret_cast<const IO*>(p.a) + row * p.E + c.e0);
    }
  };

  // Per-item metadata of this lane: the L reset flags of its steps (consecutive
  // bits of the prologue's bitmask; bits at and beyond T are zero and L divides
  // 32, so they never straddle a word) and the per-channel constants.  Loaded
  // as soon as an item is known, one iteration before it is computed, so the
  // L2 latency is off the critical path.
  unsigned meta_rs = 0;
  uint32_t meta_bx[NV], meta_ba[NV], meta_sp[NV];
  auto load_meta = [&](const Coord& c) {
    meta_rs = 0;
#pragma unroll
    for (int i = 0; i < NV; ++i) { meta_bx[i] = 0u; meta_ba[i] = 0u; meta_sp[i] = 0u; }
    if (valid_steps(c) == 0) return;
    const int ch0 = c.e0 + cv * V;
    const int t_first = (c.sc * NW + warp) * TC + seg_id * L;
    if constexpr (KIND == 2) {
      // coefficient of logical step tau is a[t+1] * ~reset[t+1], t = T-1-tau
      for (int j = 0; j < L; ++j) {
        const int tn = p.T - (t_first + j);              // t + 1
        if (tn >= 1 && tn < p.T && p.reset[(size_t)c.b * p.T + tn] != 0) meta_rs |= 1u << j;
      }
      return;
    }
    meta_rs = (p.reset_bits[(long long)c.b * p.bits_bstride + (t_first >> 5)] >> (t_first & 31)) &
              ((1u << L) - 1u);
    if constexpr (KIND == 0) {
      if constexpr (PACKED) {   // bf16 parameters are already packed pairs
        if (p.bias_x) { const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.bias_x) + ch0);
          meta_bx[0] = v.x; meta_bx[1] = v.y; meta_bx[2] = v.z; meta_bx[3] = v.w; }
        if (p.bias_a) { const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.bias_a) + ch0);
          meta_ba[0] = v.x; meta_ba[1] = v.y; meta_ba[2] = v.z; meta_ba[3] = v.w; }
        const uint4 v = *reinterpret_cast<const uint4*>(p.neg8sp_bf + ch0);
        meta_sp[0] = v.x; meta_sp[1] = v.y; meta_sp[2] = v.z; meta_sp[3] = v.w;
      } else {
        if constexpr (BF) {
          if (p.bias_x) { const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.bias_x) + ch0);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) { meta_bx[2 * i] = w[i] << 16; meta_bx[2 * i + 1] = w[i] & 0xffff0000u; } }
          if (p.bias_a) { const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.bias_a) + ch0);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) { meta_ba[2 * i] = w[i] << 16; meta_ba[2 * i + 1] = w[i] & 0xffff0000u; } }
        } else {
          if (p.bias_x) { const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.bias_x) + ch0);
            meta_bx[0] = v.x; meta_bx[1] = v.y; meta_bx[2] = v.z; meta_bx[3] = v.w; }
          if (p.bias_a) { const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.bias_a) + ch0);
            meta_ba[0] = v.x; meta_ba[1] = v.y; meta_ba[2] = v.z; meta_ba[3] = v.w; }
        }
#pragma unroll
        for (int i = 0; i < V; i += 4) {
          const uint4 v = *reinterpret_cast<const uint4*>(p.neg8sp + ch0 + i);
          meta_sp[i] = v.x; meta_sp[i + 1] = v.y; meta_sp[i + 2] = v.z; meta_sp[i + 3] = v.w;
        }
      }
    }
  };

  if (threadIdx.x == 0) {
    claim(0);
    if constexpr (STAGES == 2) claim(1);
  }
  __syncthreads();
  Coord cur = fetch(0), nxt = fetch(STAGES == 2 ? 1 : 0);
  const unsigned epoch = *p.epoch;
  issue_loads(cur, 0);
  if constexpr (STAGES == 2) issue_loads(nxt, 1);
  load_meta(cur);

  for (int it = 0; cur.item < p.nitems; ++it) {
    const int buf = STAGES == 2 ? (it & 1) : 0;
    // ticket of the item that will refill this buffer; slots 2,3 alternate so a
    // slow thread still reading the previous claim is never overwritten
    const int slot = 2 + (it & 1);
    if (threadIdx.x == 0) claim(slot);

    const int sc = cur.sc, col = cur.col, b = cur.b, e0 = cur.e0;
    const int ch0 = e0 + cv * V;
    const int nvalid = valid_steps(cur);
    const int t_first = (sc * NW + warp) * TC + seg_id * L;
    const size_t row0 = (size_t)b * p.T;
    uint4* my = stage + (size_t)buf * STAGE_U4 + (size_t)warp * L * NT * 32 + lane;
    // where the transform (kind 0: P, 1: H) of segment `sgm`, channel `ch` of the
    // tile lives: in the dead third staging slot of the lane that owns it, or
    // in the separate array.  Four consecutive channels are contiguous.
    uint4* const stage_buf = stage + (size_t)buf * STAGE_U4;
    auto xf = [&](int kind, int sgm, int ch) -> float* {
      if constexpr (ALIAS) {
        const int w = sgm / kSegs, sg = sgm % kSegs;
        const int j = kind * (V / 4) + (ch % V) / 4;
        return reinterpret_cast<float*>(stage_buf + ((w * L + j) * NT + 2) * 32 + sg * 8 + ch / V) + (ch & 3);
      } else {
        // [segment][4-channel group of the vector][lane cv][4]: the eight lanes of a
        // quarter warp (one 128-bit access each) cover 32 consecutive banks; the
        // plain [segment][channel] layout is a 2-way conflict for 8-channel vectors
        const int c = ch / V, i = ch % V;
        return (kind ? s_h : s_p) + sgm * EC + (i >> 2) * (kCvl * 4) + c * 4 + (i & 3);
      }
    };

    // warp 0 requests the state that enters this item (left by the column's
    // previous item) now, so its L2 round trip overlaps pass 1
    const int par = it & 1;
    unsigned long long pw[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) pw[i] = 0ull;
    if (warp == 0 && sc > 0) {
      const size_t src = ((size_t)(sc - 1) * p.ncols + col) * EC + lane * CPL;
#pragma unroll
      for (int i = 0; i < CPL; ++i) pw[i] = ld_relaxed_u64(p.pref + src + i);
    }

    // per-item metadata (reset bits, per-channel constants) was requested one
    // iteration ahead, see load_meta() below
    const unsigned rs_mask = meta_rs;
    uint32_t cbx[NV], cba[NV], csp[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) { cbx[i] = meta_bx[i]; cba[i] = meta_ba[i]; csp[i] = meta_sp[i]; }

    cp_async_wait<STAGES - 1>();   // this item's copies have landed (the next item's may still fly)

    // ---------------------------------------------------------- pass 1
    float P[V], H[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { P[i] = 1.0f; H[i] = 0.0f; }
    auto step1 = [&](int j, auto reset_tag) {
      constexpr bool RS = decltype(reset_tag)::value;
      const uint4 vx = my[(j * NT + 0) * 32];
      const uint4 v1 = my[(j * NT + 1) * 32];
      uint4 v2 = make_uint4(0, 0, 0, 0);
      if constexpr (KIND == 0) v2 = my[(j * NT + 2) * 32];
      const uint32_t wx[4] = {vx.x, vx.y, vx.z, vx.w};
      const uint32_t w1[4] = {v1.x, v1.y, v1.z, v1.w};
      const uint32_t w2[4] = {v2.x, v2.y, v2.z, v2.w};
      if constexpr (PACKED) {
        uint32_t sa[4], sx[4];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          uint32_t av, nx;
          if constexpr (KIND == 0) {
            gate_pair_emul<FAST, RS>(wx[i], w1[i], w2[i], cbx[i], cba[i], csp[i], av, nx);
          } else {
            av = RS ? 0u : w1[i];   // a * ~reset (:173)
            nx = wx[i];
          }
          sa[i] = av; sx[i] = nx;
          const float a0 = bf_lo(av), a1 = bf_hi(av);
          H[2 * i] = fmaf(a0, H[2 * i], bf_lo(nx));
          H[2 * i + 1] = fmaf(a1, H[2 * i + 1], bf_hi(nx));
          P[2 * i] *= a0; P[2 * i + 1] *= a1;
        }
        my[(j * NT + 0) * 32] = make_uint4(sx[0], sx[1], sx[2], sx[3]);
        my[(j * NT + 1) * 32] = make_uint4(sa[0], sa[1], sa[2], sa[3]);
      } else {
        float fx[V], f1[V], f2[V], fa[V], fn[V];
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
            gate_f32<FAST, RS>(fx[i], f1[i], f2[i], __uint_as_float(cbx[i]), __uint_as_float(cba[i]),
                               __uint_as_float(csp[i]), av, nx);
          } else {
            av = RS ? 0.0f : f1[i];
            nx = fx[i];
          }
          fa[i] = av; fn[i] = nx;
          H[i] = fmaf(av, H[i], nx);
          P[i] *= av;
        }
#pragma unroll
        for (int i = 0; i < V; i += 4) {       // x~ in slots 0.., a after them
          my[(j * NT + i / 4) * 32] = make_uint4(__float_as_uint(fn[i]), __float_as_uint(fn[i + 1]),
                                                 __float_as_uint(fn[i + 2]), __float_as_uint(fn[i + 3]));
          my[(j * NT + V / 4 + i / 4) * 32] = make_uint4(__float_as_uint(fa[i]), __float_as_uint(fa[i + 1]),
                                                         __float_as_uint(fa[i + 2]), __float_as_uint(fa[i + 3]));
        }
      }
    };
    // steps beyond T / E are identities: simply not executed.  Not unrolled:
    // one step already carries 4 independent bf16x2 chains per lane, and a
    // fully unrolled body overflows the instruction cache.
    if (rs_mask == 0u && nvalid == L) {          // common case: no document start, full segment
#pragma unroll (kP1Unroll)
      for (int j = 0; j < L; ++j) step1(j, FalseTag{});
    } else {
#pragma unroll 1
      for (int j = 0; j < nvalid; ++j) {
        if ((rs_mask >> j) & 1u) step1(j, TrueTag{});   // document start (rare): own code path
        else step1(j, FalseTag{});
      }
    }
    // ---- carries inside the warp: my segment's transform h -> P*h + H goes to
    // shared memory; each lane composes the transforms of its warp's earlier
    // segments (<= 3), the last lane group also the warp's total.
#pragma unroll
    for (int i = 0; i < V; i += 4) {
      *reinterpret_cast<float4*>(xf(0, my_seg, cv * V + i)) = make_float4(P[i], P[i + 1], P[i + 2], P[i + 3]);
      *reinterpret_cast<float4*>(xf(1, my_seg, cv * V + i)) = make_float4(H[i], H[i + 1], H[i + 2], H[i + 3]);
    }
    __syncwarp();
    float Pex[V], Hex[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { Pex[i] = 1.0f; Hex[i] = 0.0f; }
    for (int sg = 0; sg < seg_id; ++sg) {
#pragma unroll
      for (int i = 0; i < V; i += 4) {
        const float4 vp = *reinterpret_cast<const float4*>(xf(0, warp * kSegs + sg, cv * V + i));
        const float4 vh = *reinterpret_cast<const float4*>(xf(1, warp * kSegs + sg, cv * V + i));
        Hex[i] = fmaf(vp.x, Hex[i], vh.x); Hex[i + 1] = fmaf(vp.y, Hex[i + 1], vh.y);
        Hex[i + 2] = fmaf(vp.z, Hex[i + 2], vh.z); Hex[i + 3] = fmaf(vp.w, Hex[i + 3], vh.w);
        Pex[i] *= vp.x; Pex[i + 1] *= vp.y; Pex[i + 2] *= vp.z; Pex[i + 3] *= vp.w;
      }
    }
    float* wp = s_wp + (par * NW) * EC;
    float* wh = s_wh + (par * NW) * EC;
    float* c0s = s_c0 + par * EC;
    if (seg_id == kSegs - 1) {   // transform of the whole warp (chunk)
#pragma unroll
      for (int i = 0; i < V; i += 4) {
        *reinterpret_cast<float4*>(wp + warp * EC + cv * V + i) =
            make_float4(P[i] * Pex[i], P[i + 1] * Pex[i + 1], P[i + 2] * Pex[i + 2], P[i + 3] * Pex[i + 3]);
        *reinterpret_cast<float4*>(wh + warp * EC + cv * V + i) =
            make_float4(fmaf(P[i], Hex[i], H[i]), fmaf(P[i + 1], Hex[i + 1], H[i + 1]),
                        fmaf(P[i + 2], Hex[i + 2], H[i + 2]), fmaf(P[i + 3], Hex[i + 3], H[i + 3]));
      }
    }
    // ---- state entering the item.  Lane l of warp 0 owns channels
    // [l*CPL, l*CPL+CPL).  Usually the predecessor published long ago and the
    // prefetched words are valid: no second barrier is needed then.
    if (warp == 0) {
      bool ready = true;
      float c0[CPL];
      if (sc == 0) {
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
          const int ch = e0 + lane * CPL + i;
          const float* init = KIND == 2 ? p.g_last : p.h0;   // backward: the scan starts from grad(last_h)
          c0[i] = (init != nullptr && ch < p.E) ? init[(size_t)b * p.E + ch] : 0.0f;
        }
      } else {
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
          ready = ready && (unsigned)pw[i] == epoch;
          c0[i] = tagged_value(pw[i]);
        }
        ready = __all_sync(0xffffffffu, ready);
      }
#pragma unroll
      for (int i = 0; i < CPL; ++i) c0s[lane * CPL + i] = c0[i];
      if (lane == 0) s_flag[par] = ready ? 1 : 0;
    }
    __syncthreads();
    const Coord nn = fetch(slot);
    prefetch_l2(nn);
    load_meta(STAGES == 2 ? nxt : nn);   // metadata of the item computed next

    if (s_flag[par] == 0) {   // CTA-uniform slow path: decoupled look-back
      if (warp == 0) {
        float pt[CPL], ht[CPL], c0[CPL];
#pragma unroll
        for (int i = 0; i < CPL; ++i) { pt[i] = 1.0f; ht[i] = 0.0f; }
#pragma unroll
        for (int w = 0; w < NW; ++w)
#pragma unroll
          for (int i = 0; i < CPL; ++i) {
            const float pv = wp[w * EC + lane * CPL + i], hv = wh[w * EC + lane * CPL + i];
            ht[i] = fmaf(pv, ht[i], hv);
            pt[i] *= pv;
          }
        // publish my aggregate so successors can fold over it, then look back
        const size_t ws_off = (size_t)cur.item * EC + lane * CPL;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
          st_relaxed_u64(p.agg_p + ws_off + i, pack_tagged(pt[i], epoch));
          st_relaxed_u64(p.agg_h + ws_off + i, pack_tagged(ht[i], epoch));
        }
        int j = sc - 1;
        for (;;) {
          const size_t src = ((size_t)j * p.ncols + col) * EC + lane * CPL;
          bool ready = true;
#pragma unroll
          for (int i = 0; i < CPL; ++i) {
            pw[i] = ld_relaxed_u64(p.pref + src + i);
            ready = ready && (unsigned)pw[i] == epoch;
          }
          if (__all_sync(0xffffffffu, ready)) break;
          bool agg = true;
#pragma unroll
          for (int i = 0; i < CPL; ++i) {
            agg = agg && (unsigned)ld_relaxed_u64(p.agg_p + src + i) == epoch &&
                  (unsigned)ld_relaxed_u64(p.agg_h + src + i) == epoch;
          }
          if (__all_sync(0xffffffffu, agg)) { --j; continue; }   // item 0 of a column always publishes its state
          __nanosleep(20);
        }
#pragma unroll
        for (int i = 0; i < CPL; ++i) c0[i] = tagged_value(pw[i]);
        for (int k = j + 1; k < sc; ++k) {   // fold the aggregates seen on the way, left to right
          const size_t off = ((size_t)k * p.ncols + col) * EC + lane * CPL;
#pragma unroll
          for (int i = 0; i < CPL; ++i)
            c0[i] = fmaf(tagged_value(ld_relaxed_u64(p.agg_p + off + i)), c0[i],
                         tagged_value(ld_relaxed_u64(p.agg_h + off + i)));
        }
#pragma unroll
        for (int i = 0; i < CPL; ++i) c0s[lane * CPL + i] = c0[i];
      }
      __syncthreads();
    }

    // ---- the last warp publishes the state leaving the item: the item's
    // aggregate (warp transforms composed in order) applied to the carry-in --
    // the same expression the look-back uses, so values never depend on timing.
    if (warp == NW - 1 && sc + 1 < p.nchunks) {
      const size_t ws_off = (size_t)cur.item * EC + lane * CPL;
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        float pt = 1.0f, ht = 0.0f;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          const float pv = wp[w * EC + lane * CPL + i], hv = wh[w * EC + lane * CPL + i];
          ht = fmaf(pv, ht, hv);
          pt *= pv;
        }
        st_relaxed_u64(p.pref + ws_off + i, pack_tagged(fmaf(pt, c0s[lane * CPL + i], ht), epoch));
      }
    }

    // ---- carry into my segment: fold the item's earlier chunks, then my
    // warp's earlier segments
    float h[V];
#pragma unroll
    for (int i = 0; i < V; i += 4) {
      const float4 vc = *reinterpret_cast<const float4*>(c0s + cv * V + i);
      h[i] = vc.x; h[i + 1] = vc.y; h[i + 2] = vc.z; h[i + 3] = vc.w;
    }
    for (int w = 0; w < warp; ++w) {
#pragma unroll
      for (int i = 0; i < V; i += 4) {
        const float4 vp = *reinterpret_cast<const float4*>(wp + w * EC + cv * V + i);
        const float4 vh = *reinterpret_cast<const float4*>(wh + w * EC + cv * V + i);
        h[i] = fmaf(vp.x, h[i], vh.x); h[i + 1] = fmaf(vp.y, h[i + 1], vh.y);
        h[i + 2] = fmaf(vp.z, h[i + 2], vh.z); h[i + 3] = fmaf(vp.w, h[i + 3], vh.w);
      }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) h[i] = fmaf(Pex[i], h[i], Hex[i]);

    // ---------------------------------------------------------- pass 2 (replay)
    IO* yrow = reinterpret_cast<IO*>(p.y) + (row0 + t_first) * p.E + ch0;
    // KIND 2: grad of a for the physical step t: dh_t * h_{t-1} (h_{-1} = h0),
    // zero at a document start (a was multiplied by ~reset, :173)
    auto grad_a = [&](int j, const float (&dh)[V], const uint4& vhp) -> uint4 {
      const int t = p.T - 1 - (t_first + j);
      float hp[V];
      if (t == 0) {
#pragma unroll
        for (int i = 0; i < V; ++i) hp[i] = p.h0 != nullptr ? p.h0[(size_t)b * p.E + ch0 + i] : 0.0f;
      } else if constexpr (BF) {
        const uint32_t w[4] = {vhp.x, vhp.y, vhp.z, vhp.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { hp[2 * i] = bf_lo(w[i]); hp[2 * i + 1] = bf_hi(w[i]); }
      } else {
        hp[0] = __uint_as_float(vhp.x); hp[1] = __uint_as_float(vhp.y);
        hp[2] = __uint_as_float(vhp.z); hp[3] = __uint_as_float(vhp.w);
      }
      float g[V];
      const bool rz = p.reset[row0 + t] != 0;
#pragma unroll
      for (int i = 0; i < V; ++i) g[i] = rz ? 0.0f : __fmul_rn(dh[i], hp[i]);
      if constexpr (BF)
        return make_uint4(pack_bf2(g[0], g[1]), pack_bf2(g[2], g[3]), pack_bf2(g[4 % V], g[5 % V]),
                          pack_bf2(g[6 % V], g[7 % V]));
      else
        return make_uint4(__float_as_uint(g[0]), __float_as_uint(g[1]), __float_as_uint(g[2]),
                          __float_as_uint(g[3]));
    };
    auto step2 = [&](int j) {
      uint4 out;
      if constexpr (PACKED) {
        const uint4 vn = my[(j * NT + 0) * 32];
        const uint4 va = my[(j * NT + 1) * 32];
        const uint32_t sx[4] = {vn.x, vn.y, vn.z, vn.w}, sa[4] = {va.x, va.y, va.z, va.w};
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          if constexpr (FAST) {
            h[2 * i] = fmaf(bf_lo(sa[i]), h[2 * i], bf_lo(sx[i]));
            h[2 * i + 1] = fmaf(bf_hi(sa[i]), h[2 * i + 1], bf_hi(sx[i]));
          } else {   // mul then add, as the reference loop (:196)
            h[2 * i] = __fadd_rn(__fmul_rn(bf_lo(sa[i]), h[2 * i]), bf_lo(sx[i]));
            h[2 * i + 1] = __fadd_rn(__fmul_rn(bf_hi(sa[i]), h[2 * i + 1]), bf_hi(sx[i]));
          }
          o[i] = pack_bf2(h[2 * i], h[2 * i + 1]);
        }
        out = make_uint4(o[0], o[1], o[2], o[3]);
      } else {
        float fn[V], fa[V];
#pragma unroll
        for (int i = 0; i < V; i += 4) {
          const uint4 vn = my[(j * NT + i / 4) * 32];
          const uint4 va = my[(j * NT + V / 4 + i / 4) * 32];
          fn[i] = __uint_as_float(vn.x); fn[i + 1] = __uint_as_float(vn.y);
          fn[i + 2] = __uint_as_float(vn.z); fn[i + 3] = __uint_as_float(vn.w);
          fa[i] = __uint_as_float(va.x); fa[i + 1] = __uint_as_float(va.y);
          fa[i + 2] = __uint_as_float(va.z); fa[i + 3] = __uint_as_float(va.w);
        }
#pragma unroll
        for (int i = 0; i < V; ++i) {
          if constexpr (FAST) h[i] = fmaf(fa[i], h[i], fn[i]);
          else h[i] = __fadd_rn(__fmul_rn(fa[i], h[i]), fn[i]);
        }
        if constexpr (BF) {
          out = make_uint4(pack_bf2(h[0], h[1]), pack_bf2(h[2], h[3]),
                           pack_bf2(h[4 % V], h[5 % V]), pack_bf2(h[6 % V], h[7 % V]));
        } else {
          out = make_uint4(__float_as_uint(h[0]), __float_as_uint(h[1]),
                           __float_as_uint(h[2]), __float_as_uint(h[3]));
        }
      }
      if constexpr (KIND == 2) {
        const size_t off = ((row0 + p.T - 1 - (t_first + j)) * p.E + ch0);
        stg_stream(reinterpret_cast<IO*>(p.y) + off, out);
        stg_stream(reinterpret_cast<IO*>(p.da) + off, grad_a(j, h, my[(j * NT + 2) * 32]));
      } else {
        stg_stream(yrow + (size_t)j * p.E, out);
      }
    };
#pragma unroll 2
    for (int j = 0; j < nvalid; ++j) step2(j);
    // hidden state after the last valid step (padding steps are identities)
    if (p.last_h != nullptr && sc == p.nchunks - 1 && my_seg == NSEG - 1 && ch0 < p.E) {
      if constexpr (KIND == 2) {
        // grad of h0 = a_0 * ~reset_0 * dh_0 (h is dh_0 here: the scan ended at t = 0)
        const bool rz = p.reset[row0] != 0;
#pragma unroll
        for (int i = 0; i < V; ++i) {
          float a0;
          if constexpr (BF) a0 = __uint_as_float((uint32_t)reinterpret_cast<const uint16_t*>(p.a)[row0 * p.E + ch0 + i] << 16);
          else a0 = reinterpret_cast<const float*>(p.a)[row0 * p.E + ch0 + i];
          h[i] = rz ? 0.0f : __fmul_rn(a0, h[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < V; i += 4)
        *reinterpret_cast<float4*>(p.last_h + (size_t)b * p.E + ch0 + i) =
            make_float4(h[i], h[i + 1], h[i + 2], h[i + 3]);
    }

    // refill this buffer with the newly claimed item (own slots only)
    issue_loads(nn, buf);
    if constexpr (STAGES == 2) { cur = nxt; nxt = nn; } else { cur = nn; }
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------
// Per-call prologue (one small launch):
//   * resets the ticket and opens a new epoch for the exchange words;
//   * evaluates -8 * softplus(a_param) once per channel with accurate libdevice
//     math, as fp32 and as bf16.  emulate != 0 (bf16 reference mode): softplus
//     is rounded to bf16 first (layers.py:352), so the bf16 copy is exact;
//   * turns `segment_pos == 0` (layers.py:345) or the rnn_scan reset mask into
//     a bitmask, one 32-bit word per 32 time steps.
// ---------------------------------------------------------------------------
struct PrologueParams {
  const void* a_param;        // [E] or null (plain scan)
  float* neg8sp; uint16_t* neg8sp_bf;
  int E, is_bf16, emulate;
  int* counter; unsigned* epoch;   // null for the strict kernels
  const void* seg; int seg_is_i64; long long seg_bstride;   // RG-LRU
  const unsigned char* reset;                               // plain scan
  unsigned* reset_bits; int rows, T, words_per_row;
};

__global__ void scan_prologue_kernel(const PrologueParams q) {
  // programmatic dependent launch: a consumer launched with the PDL attribute
  // (the fused RG-LRU kernel) may start its own set-up now; it orders itself
  // after this grid's writes with griddepcontrol.wait.  No-op otherwise.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && q.counter != nullptr) { *q.counter = 0; *q.epoch = *q.epoch + 1u; }
  // per-head ticket counters of the fused kernel's dynamic schedule: words 8 .. 63 of the scratch header
  if (i >= 8 && i < 64 && q.counter != nullptr) q.counter[i] = 0;
  if (q.a_param != nullptr && i < q.E) {
    float ap = q.is_bf16 ? __uint_as_float((uint32_t)reinterpret_cast<const uint16_t*>(q.a_param)[i] << 16)
                         : reinterpret_cast<const float*>(q.a_param)[i];
    float sp = softplus_f(ap);
    if (q.is_bf16 && q.emulate) sp = round_bf(sp);
    const float v = -8.0f * sp;
    q.neg8sp[i] = v;
    q.neg8sp_bf[i] = (uint16_t)(pack_bf2(v, v) & 0xffffu);
  }
  // one thread per (row, step); a warp covers the 32 steps of one bitmask word
  const int tpad = q.words_per_row * 32;
  if (q.reset_bits != nullptr && i < q.rows * tpad) {
    const int r = i / tpad, t = i - r * tpad;
    bool rs = false;
    if (t < q.T) {
      if (q.seg != nullptr) rs = seg_is_zero(q.seg, q.seg_is_i64 != 0, (long long)r * q.seg_bstride + t);
      else rs = q.reset[(size_t)r * q.T + t] != 0;
    }
    const unsigned bits = __ballot_sync(0xffffffffu, rs);
    if ((threadIdx.x & 31) == 0) q.reset_bits[r * q.words_per_row + (t >> 5)] = bits;
  }
  // When launched as a programmatic dependent (fused path) this grid may have run
  // under the tail of its predecessor (the Conv1D kernel, whose output it does
  // not read).  It must not COMPLETE before the predecessor has: the fused kernel
  // takes "prologue complete" to mean "x complete".  No-op for a plain launch.
  asm volatile("griddepcontrol.wait;" ::: "memory");
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
      const float g1 = load_io<IO>(p.gemm_x, row * p.gate_ld + gate_off(e, p.gate_bw));
      const float g2 = load_io<IO>(p.gemm_a, row * p.gate_ld + gate_off(e, p.gate_bw));
      if constexpr (EMUL) {
        // scalar spelling of gate_pair_emul: round after every eager op
        float px = round_bf(g1 + bx), pa = round_bf(g2 + ba);
        if constexpr (FAST) { px = fmaxf(px, -43.0f); pa = fmaxf(pa, -43.0f); }
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
        if (rs) gate_f32<FAST, true>(xv, g1, g2, bx, ba, sp8, av, nx);
        else gate_f32<FAST, false>(xv, g1, g2, bx, ba, sp8, av, nx);
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

// ---------------------------------------------------------------------------
// Decode step (T == 1): rnn_scan's sampling branch (layers.py:175-182) fused with
// the gate math.  One thread per 16-byte channel vector, ONE launch (softplus is
// evaluated inline: B*E elements only).  y = r(a*h0 + x~) with separate fp32
// mul and add; last_h is the unrounded fp32 value; h0 == nullptr returns x~.
// ---------------------------------------------------------------------------
template <typename IO, int ARITH>
__global__ void __launch_bounds__(128)
rglru_step_kernel(const ScanParams p, const void* a_param) {
  using Tr = ScanTraits<IO, 0, ARITH>;
  constexpr int V = Tr::V;
  constexpr bool BF = Tr::BF;
  constexpr bool FAST = Tr::FAST;
  constexpr bool PACKED = Tr::PACKED;
  const int vecs = p.E / V;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.B * vecs) return;
  const int b = idx / vecs, ch0 = (idx - b * vecs) * V;
  const bool rs = seg_is_zero(p.seg, p.seg_is_i64 != 0, (long long)b * p.seg_bstride);
  const size_t row = b;   // T == 1
  const uint4 vx = ldg_stream(reinterpret_cast<const IO*>(p.x) + row * p.E + ch0);
  const int g0 = gate_off(ch0, p.gate_bw);
  const uint4 v1 = ldg_stream(reinterpret_cast<const IO*>(p.gemm_x) + row * p.gate_ld + g0);
  const uint4 v2 = ldg_stream(reinterpret_cast<const IO*>(p.gemm_a) + row * p.gate_ld + g0);
  float av[V], nx[V], bx[V], ba[V], sp8[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float ap = load_io<IO>(a_param, ch0 + i);
    float sp = softplus_f(ap);
    if (PACKED) sp = round_bf(sp);
    sp8[i] = -8.0f * sp;
    bx[i] = p.bias_x ? load_io<IO>(p.bias_x, ch0 + i) : 0.0f;
    ba[i] = p.bias_a ? load_io<IO>(p.bias_a, ch0 + i) : 0.0f;
  }
  const uint32_t wx[4] = {vx.x, vx.y, vx.z, vx.w};
  const uint32_t w1[4] = {v1.x, v1.y, v1.z, v1.w};
  const uint32_t w2[4] = {v2.x, v2.y, v2.z, v2.w};
  if constexpr (PACKED) {
#pragma unroll
    for (int i = 0; i < V / 2; ++i) {
      uint32_t a2, n2;
      const uint32_t cbx = pack_bf2(bx[2 * i], bx[2 * i + 1]), cba = pack_bf2(ba[2 * i], ba[2 * i + 1]);
      const uint32_t csp = pack_bf2(sp8[2 * i], sp8[2 * i + 1]);
      if (rs) gate_pair_emul<FAST, true>(wx[i], w1[i], w2[i], cbx, cba, csp, a2, n2);
      else gate_pair_emul<FAST, false>(wx[i], w1[i], w2[i], cbx, cba, csp, a2, n2);
      av[2 * i] = bf_lo(a2); av[2 * i + 1] = bf_hi(a2);
      nx[2 * i] = bf_lo(n2); nx[2 * i + 1] = bf_hi(n2);
    }
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float fx, f1, f2;
      if constexpr (BF) {
        fx = (i & 1) ? bf_hi(wx[i / 2]) : bf_lo(wx[i / 2]);
        f1 = (i & 1) ? bf_hi(w1[i / 2]) : bf_lo(w1[i / 2]);
        f2 = (i & 1) ? bf_hi(w2[i / 2]) : bf_lo(w2[i / 2]);
      } else {
        fx = __uint_as_float(wx[i % 4]); f1 = __uint_as_float(w1[i % 4]); f2 = __uint_as_float(w2[i % 4]);
      }
      if (rs) gate_f32<FAST, true>(fx, f1, f2, bx[i], ba[i], sp8[i], av[i], nx[i]);
      else gate_f32<FAST, false>(fx, f1, f2, bx[i], ba[i], sp8[i], av[i], nx[i]);
    }
  }
  float h[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    if (p.h0 == nullptr) h[i] = nx[i];                                           // :177-178
    else h[i] = __fadd_rn(__fmul_rn(av[i], p.h0[(size_t)b * p.E + ch0 + i]), nx[i]);   // :181
  }
  uint4 out;
  if constexpr (BF) {
    out = make_uint4(pack_bf2(h[0], h[1]), pack_bf2(h[2], h[3]), pack_bf2(h[4 % V], h[5 % V]),
                     pack_bf2(h[6 % V], h[7 % V]));
  } else {
    out = make_uint4(__float_as_uint(h[0]), __float_as_uint(h[1]), __float_as_uint(h[2]),
                     __float_as_uint(h[3]));
  }
  stg_stream(reinterpret_cast<IO*>(p.y) + row * p.E + ch0, out);
  if (p.last_h != nullptr) {
    // (without h0 the reference returns x[:, 0] widened; in reference mode x~ is
    // already bf16-exact, so the fp32 value below is that very number)
#pragma unroll
    for (int i = 0; i < V; i += 4)
      *reinterpret_cast<float4*>(p.last_h + (size_t)b * p.E + ch0 + i) =
          make_float4(h[i], h[i + 1], h[i + 2], h[i + 3]);
  }
}

}  // namespace cg
