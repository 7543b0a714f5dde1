// Fused RG-LRU for sm_100a: block-diagonal gate GEMMs on tcgen05 tensor cores
// with TMEM accumulators, gate math and the chunked linear-recurrence scan in
// the epilogue -- the gate pre-activations never reach HBM.
//
// Replaces, in ONE kernel, reference recurrentgemma/torch/layers.py:345-375
// (RGLRU.forward) INCLUDING the two BlockDiagonalLinear GEMMs (:133-142, :348-349)
// and rnn_scan (:146-199).  SURVEY.md section 8(f) row F1.
//
// Formulation (transposed GEMM, channels on TMEM lanes, time on TMEM columns):
//   work column = (batch row b, 128 channels of one head);  tile = 64 time steps
//   of one column, issued as two "granules" of 32 steps.  Per granule the MMA
//   warp issues  D_x = Wx^T . X^T,  D_a = Wa^T . X^T  (M = 128 channels, N = 32
//   steps, K = head width) and  D_t = I . X^T  (identity: the tensor core
//   transposes the activations, exactly) into a 96-column TMEM slot.
//   An epilogue thread owns ONE channel: tcgen05.ld hands it the pre-activations
//   and x of consecutive time steps in registers, so the gate math works on
//   bf16x2 pairs of (t, t+1), the recurrence is a register-resident sequential
//   scan, and there is no shared-memory staging of activations at all.
//   Tiles of one column are chained through global memory with the decoupled
//   look-back of cg_scan.cuh (tagged 64-bit words, left-to-right folds only =>
//   bit-reproducible).
//
// Roles (320 threads, 1 CTA / SM, persistent):
//   warps 0-3, 4-7  two epilogue warpgroups (tiles alternate between them; each
//                   owns two TMEM slots, so the MMAs of its next tile run while
//                   it is still working)
//   warp 8          TMA producer: gate weights (packed, pre-swizzled) once per
//                   column family, then one X tile [64 steps x head width] per
//                   tile through a 3-D tensor map (SWIZZLE_128B, zero fill
//                   beyond T)
//   warp 9          MMA issuer (one thread), tcgen05.commit -> mbarriers
// CTA i works on column family (head, channel half) = i % families; the CTAs
// of one family take its tiles round-robin in time-major order, so look-back
// dependencies always point to tiles that are already running (all CTAs are
// co-resident: grid <= #SMs, 1 CTA / SM).
#pragma once

#include <cuda.h>

#include "cg_common.cuh"
#include "cg_scan.cuh"

namespace cg {
namespace fused {

constexpr int kGran = 32;        // time steps per MMA granule (UMMA N)
constexpr int kMch = 128;        // channels per work column (UMMA M)
constexpr int kSlotCols = 96;    // TMEM columns per granule: pre_x | pre_a | x^T
constexpr int kSlots = 4;        // TMEM slots: 2 per epilogue warpgroup
constexpr int kTmemCols = 512;
constexpr int kEpiWarps = 8;
constexpr int kThreads = (kEpiWarps + 2) * 32;
constexpr uint32_t kKBlockBytes = 128u * 128u;   // 128 rows x 64 bf16 (one swizzle-128B K block)

struct FusedParams {
  const unsigned char* wpack;      // [families][2 gates][KB][128 rows][128 B], swizzled (pack kernel)
  const unsigned char* ident;      // [2][128 rows][128 B] identity, swizzled
  const uint16_t* bias_x;          // [E] or null
  const uint16_t* bias_a;
  const uint16_t* neg8sp_bf;       // [E] -8*softplus(a_param) as bf16 (prologue)
  const unsigned* reset_bits;      // [rows][words] bit t%32 of word t/32
  long long bits_bstride;          // words per batch row (0 = broadcast)
  int words;                       // words per row
  const float* h0;                 // [B,E] or null
  uint16_t* y;                     // [B,T,E]
  float* last_h;                   // [B,E] or null
  const unsigned* epoch;
  unsigned long long* agg_p;       // [families][ntt][B][128]
  unsigned long long* agg_h;
  unsigned long long* pref;
  uint16_t* dbg;                   // optional [3][B][T][E]: rounded pre_x, pre_a, x^T
  int* err;                        // watchdog flag
  int B, T, E;
  int ntt;                         // time tiles per batch row
  int families;                    // E / 128
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must end the launch, never hang the device.  On
// a timeout the wait raises the error flag (workspace header) and returns; once
// the flag is up every later wait returns at once, so the kernel drains (with
// garbage results) and the host finds the flag.
constexpr long long kWatchdogCycles = 1000000000LL;   // ~0.5 s
__device__ __forceinline__ bool watchdog_expired(long long t0, int* err, int code, unsigned& polls) {
  if ((++polls & 1023u) != 0u) return false;
  if (*reinterpret_cast<volatile int*>(err) != 0) return true;
  if (clock64() - t0 > kWatchdogCycles) { atomicCAS(err, 0, code); return true; }
  return false;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err, int code) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  unsigned polls = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (watchdog_expired(t0, err, code, polls)) return;
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :: "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      :: "r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t holder_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
               :: "r"(holder_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, bf16 inputs, fp32 accumulate, one CTA.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               :: "r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets lane
// (quadrant base + i), registers = columns.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      " {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_u16(uint16_t* p, uint32_t v) {
  asm volatile("st.global.u16 [%0], %1;" :: "l"(p), "h"(static_cast<uint16_t>(v)) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start address >> 4 | LBO (ignored for swizzled K-major, 1) | SBO = 1024 B between
// 8-row groups | version 1 (Blackwell) | layout type 2 (128 B swizzle).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return static_cast<uint64_t>((smem_addr >> 4) & 0x3fffu) | (1ull << 16) | (64ull << 32) |
         (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, bf16 A
// and B, both K-major, N at bits 17.., M at bits 24..
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

template <int KB, int GPT>
struct FusedCfg {
  static constexpr int kTileT = kGran * GPT;
  static constexpr uint32_t kWBytes = 2u * KB * kKBlockBytes;
  static constexpr uint32_t kIBytes = 2u * kKBlockBytes;
  static constexpr uint32_t kXStageBytes = static_cast<uint32_t>(KB) * kTileT * 128u;
  static constexpr int kXStages = 65536 / kXStageBytes;
  static constexpr int kBars = 2 + 2 * kXStages + 2 * kSlots;
  static constexpr size_t kSmemBytes = 1024 + kWBytes + kIBytes + kXStages * kXStageBytes + kBars * 8 + 16;
  static_assert(KB % 2 == 0, "head width must be a multiple of 128");
  static_assert(kXStages >= 2, "need at least two X stages");
};

// ---------------------------------------------------------------------------
// One-time packing of the gate weights into the shared-memory image the MMA
// reads (K-major rows of 64 bf16 = 128 B, 16-byte chunks XOR-swizzled with the
// row index, exactly what a SWIZZLE_128B TMA box would have produced):
//   wpack[family][gate][kb][row r][k]  =  w_gate[head][kb*64 + k][cb*128 + r]
// (reference weight layout [H, bw_in, bw_out], layers.py:103; y = x @ w[h]).
// One thread per 16-byte chunk.  Also writes the identity operand.
// ---------------------------------------------------------------------------
__global__ void pack_gate_weights_kernel(const uint16_t* __restrict__ wx, const uint16_t* __restrict__ wa,
                                         unsigned char* __restrict__ wpack, unsigned char* __restrict__ ident,
                                         int H, int bw) {
  const int kb_per_head = bw / 64;
  const int cbs = bw / 128;
  const int families = H * cbs;
  const long long chunks_w = (long long)families * 2 * kb_per_head * 128 * 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < chunks_w) {
    const int cidx = idx & 7;
    const int r = (idx >> 3) & 127;
    long long rest = idx >> 10;
    const int kb = rest % kb_per_head; rest /= kb_per_head;
    const int gate = rest & 1; rest >>= 1;
    const int fam = (int)rest;
    const int head = fam / cbs, cb = fam % cbs;
    const uint16_t* w = (gate == 0 ? wx : wa) + (size_t)head * bw * bw;
    uint16_t v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = w[(size_t)(kb * 64 + cidx * 8 + i) * bw + cb * 128 + r];
    const size_t off = ((((size_t)fam * 2 + gate) * kb_per_head + kb) * 128 + r) * 128 +
                       (size_t)((cidx ^ (r & 7)) << 4);
    uint4 o;
    o.x = v[0] | ((uint32_t)v[1] << 16); o.y = v[2] | ((uint32_t)v[3] << 16);
    o.z = v[4] | ((uint32_t)v[5] << 16); o.w = v[6] | ((uint32_t)v[7] << 16);
    *reinterpret_cast<uint4*>(wpack + off) = o;
  } else if (idx < chunks_w + 2 * 128 * 8) {
    const int j = (int)(idx - chunks_w);
    const int cidx = j & 7, r = (j >> 3) & 127, kb = j >> 10;
    uint16_t v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (kb * 64 + cidx * 8 + i == r) ? 0x3f80 : 0;
    const size_t off = ((size_t)kb * 128 + r) * 128 + (size_t)((cidx ^ (r & 7)) << 4);
    uint4 o;
    o.x = v[0] | ((uint32_t)v[1] << 16); o.y = v[2] | ((uint32_t)v[3] << 16);
    o.z = v[4] | ((uint32_t)v[5] << 16); o.w = v[6] | ((uint32_t)v[7] << 16);
    *reinterpret_cast<uint4*>(ident + off) = o;
  }
}

// ---------------------------------------------------------------------------
// The fused kernel.  KB = head width / 64 (K blocks), GPT = granules per tile.
// ---------------------------------------------------------------------------
template <int KB, int GPT, bool FAST, bool DBG>
__global__ void __launch_bounds__(kThreads, 1)
rglru_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const FusedParams p) {
  using Cfg = FusedCfg<KB, GPT>;
  constexpr int TILE_T = Cfg::kTileT;
  constexpr int NP = TILE_T / 2;            // bf16x2 pairs of consecutive steps per tile
  constexpr int CBS = KB / 2;               // 128-channel halves per head
  constexpr int XS = Cfg::kXStages;
  constexpr uint32_t IDESC = umma_idesc(kMch, kGran);

  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  unsigned char* sm = smem_raw + (base - raw_addr);
  const uint32_t sW = base, sI = sW + Cfg::kWBytes, sX = sI + Cfg::kIBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + Cfg::kWBytes + Cfg::kIBytes + XS * Cfg::kXStageBytes);
  uint64_t* w_full = bars;
  uint64_t* w_empty = bars + 1;
  uint64_t* x_full = bars + 2;
  uint64_t* x_empty = x_full + XS;
  uint64_t* t_full = x_empty + XS;
  uint64_t* t_empty = t_full + kSlots;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(t_empty + kSlots);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    mbar_init(w_empty, 1);
    for (int i = 0; i < XS; ++i) { mbar_init(x_full + i, 1); mbar_init(x_empty + i, 1); }
    for (int i = 0; i < kSlots; ++i) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, 4); }
    fence_mbar_init();
  }
  if (warp == kEpiWarps) tmem_alloc(smem_u32(tmem_holder), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  // column-family schedule of this CTA (identical in every role)
  const int G = gridDim.x, nfam = p.families;
  const bool spread = G >= nfam;                  // several CTAs share one family
  const int fam_step = spread ? nfam : G;
  const int rank = spread ? blockIdx.x / nfam : 0;
  const int ntiles = p.ntt * p.B;                 // tiles of one family, time-major: ticket = tt*B + b
  auto family_ctas = [&](int fam) { return spread ? (G - 1 - fam) / nfam + 1 : 1; };
  auto my_tiles = [&](int fam) {
    const int nc = family_ctas(fam);
    return rank < ntiles ? (ntiles - rank + nc - 1) / nc : 0;
  };

  if (warp == kEpiWarps) {
    // ===================================================== TMA producer
    if (lane == 0) {
      uint32_t xq = 0, witer = 0;
      for (int fam = blockIdx.x % nfam; fam < nfam; fam += fam_step) {
        const int nmine = my_tiles(fam);
        if (nmine == 0) continue;
        const int nc = family_ctas(fam);
        if (witer > 0) mbar_wait(w_empty, (witer - 1) & 1, p.err, 1);
        mbar_expect_tx(w_full, Cfg::kWBytes + Cfg::kIBytes);
        const unsigned char* wsrc = p.wpack + (size_t)fam * Cfg::kWBytes;
#pragma unroll 1
        for (uint32_t off = 0; off < Cfg::kWBytes; off += kKBlockBytes)
          bulk_load(sW + off, wsrc + off, kKBlockBytes, w_full);
#pragma unroll 1
        for (uint32_t off = 0; off < Cfg::kIBytes; off += kKBlockBytes)
          bulk_load(sI + off, p.ident + off, kKBlockBytes, w_full);
        ++witer;
        const int c_head = (fam / CBS) * (KB * 64);
#pragma unroll 1
        for (int n = 0; n < nmine; ++n) {
          const int ticket = rank + n * nc;
          const int tt = ticket / p.B, b = ticket - tt * p.B;
          const uint32_t stage = xq % XS, use = xq / XS;
          mbar_wait(x_empty + stage, (use & 1) ^ 1, p.err, 2);
          mbar_expect_tx(x_full + stage, Cfg::kXStageBytes);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb)
            tma_load_3d(sX + stage * Cfg::kXStageBytes + kb * (TILE_T * 128), &tmap_x, x_full + stage,
                        c_head + kb * 64, tt * TILE_T, b);
          ++xq;
        }
      }
    }
    __syncwarp();
  } else if (warp == kEpiWarps + 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      uint32_t xq = 0, witer = 0, q = 0, gq0 = 0, gq1 = 0;
      for (int fam = blockIdx.x % nfam; fam < nfam; fam += fam_step) {
        const int nmine = my_tiles(fam);
        if (nmine == 0) continue;
        const int cb = fam % CBS;
        mbar_wait(w_full, witer & 1, p.err, 3);
        tc_fence_after();
#pragma unroll 1
        for (int n = 0; n < nmine; ++n) {
          const uint32_t wg = q & 1;
          const uint32_t stage = xq % XS;
          mbar_wait(x_full + stage, (xq / XS) & 1, p.err, 4);
          tc_fence_after();
          const uint32_t xs_addr = sX + stage * Cfg::kXStageBytes;
#pragma unroll 1
          for (int g = 0; g < GPT; ++g) {
            const uint32_t gq = wg ? gq1 : gq0;
            const uint32_t slot = wg * 2 + (gq & 1);
            mbar_wait(t_empty + slot, ((gq >> 1) & 1) ^ 1, p.err, 5);
            tc_fence_after();
            const uint32_t dcol = tmem_base + slot * kSlotCols;
            const uint32_t xrow = xs_addr + g * (kGran * 128);   // rows of this granule inside each K block
#pragma unroll
            for (int gate = 0; gate < 2; ++gate) {
#pragma unroll
              for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma_bf16(dcol + gate * kGran,
                            umma_desc(sW + (gate * KB + kb) * kKBlockBytes + k * 32),
                            umma_desc(xrow + kb * (TILE_T * 128) + k * 32), IDESC, (kb | k) != 0);
                }
              }
            }
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_bf16(dcol + 2 * kGran, umma_desc(sI + kb * kKBlockBytes + k * 32),
                          umma_desc(xrow + (2 * cb + kb) * (TILE_T * 128) + k * 32), IDESC, (kb | k) != 0);
              }
            }
            umma_commit(t_full + slot);
            if (wg) ++gq1; else ++gq0;
          }
          umma_commit(x_empty + stage);
          ++xq; ++q;
        }
        umma_commit(w_empty);
        ++witer;
      }
    }
    __syncwarp();
  } else {
    // ===================================================== epilogue warpgroups
    const int wg = warp >> 2;
    const int chl = (warp & 3) * 32 + lane;           // TMEM lane = channel inside the column
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const unsigned epoch = *p.epoch;
    uint32_t q = 0, gq = 0;
    for (int fam = blockIdx.x % nfam; fam < nfam; fam += fam_step) {
      const int nmine = my_tiles(fam);
      if (nmine == 0) continue;
      const int nc = family_ctas(fam);
      const int ch = fam * kMch + chl;
      uint32_t bx2 = 0, ba2 = 0;
      if (p.bias_x != nullptr) { const uint32_t v = p.bias_x[ch]; bx2 = v | (v << 16); }
      if (p.bias_a != nullptr) { const uint32_t v = p.bias_a[ch]; ba2 = v | (v << 16); }
      uint32_t sp2 = p.neg8sp_bf[ch]; sp2 |= sp2 << 16;
#pragma unroll 1
      for (int n = 0; n < nmine; ++n, ++q) {
        if ((q & 1u) != static_cast<uint32_t>(wg)) continue;
        const int ticket = rank + n * nc;
        const int tt = ticket / p.B, b = ticket - tt * p.B;
        const int t0 = tt * TILE_T;
        uint32_t A2[NP], X2[NP];
        float P = 1.0f, Hh = 0.0f;
        // ------------------------------------------------------ pass 1
#pragma unroll
        for (int g = 0; g < GPT; ++g) {
          const uint32_t slot = wg * 2 + (gq & 1);
          mbar_wait(t_full + slot, (gq >> 1) & 1, p.err, 6);
          tc_fence_after();
          const int word = (t0 >> 5) + g;
          const unsigned rbits = word < p.words ? p.reset_bits[(long long)b * p.bits_bstride + word] : 0u;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t dx[16], da[16], dt[16];
            const uint32_t col = tmem_base + lane_base + slot * kSlotCols + c * 16;
            tmem_ld16(col, dx);
            tmem_ld16(col + kGran, da);
            tmem_ld16(col + 2 * kGran, dt);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int pi = g * 16 + c * 8 + i;           // pair index inside the tile
              const int tl = t0 + 2 * pi;                  // time of the low half
              // the GEMM output the reference materialises in bf16 (:136-142)
              const uint32_t gxr = pack_bf2(__uint_as_float(dx[2 * i]), __uint_as_float(dx[2 * i + 1]));
              const uint32_t gar = pack_bf2(__uint_as_float(da[2 * i]), __uint_as_float(da[2 * i + 1]));
              const uint32_t xc = pack_bf2(__uint_as_float(dt[2 * i]), __uint_as_float(dt[2 * i + 1]));
              uint32_t a2, n2;
              gate_pair_emul<FAST, false>(xc, gxr, gar, bx2, ba2, sp2, a2, n2);
              const unsigned r2 = (rbits >> (c * 16 + 2 * i)) & 3u;
              if (r2 != 0u) {                              // document start inside the pair (rare, warp-uniform)
                uint32_t az, nr;
                gate_pair_emul<FAST, true>(xc, gxr, gar, bx2, ba2, sp2, az, nr);
                if (r2 & 1u) { a2 &= 0xffff0000u; n2 = (n2 & 0xffff0000u) | (nr & 0x0000ffffu); }
                if (r2 & 2u) { a2 &= 0x0000ffffu; n2 = (n2 & 0x0000ffffu) | (nr & 0xffff0000u); }
              }
              if (tl + 1 >= p.T) {                         // steps beyond T are identities
                if (tl >= p.T) { a2 = kOne2; n2 = 0u; }
                else { a2 = (a2 & 0x0000ffffu) | 0x3f800000u; n2 &= 0x0000ffffu; }
              }
              if constexpr (DBG) {
                const size_t plane = (size_t)p.B * p.T * p.E;
                const size_t o0 = ((size_t)b * p.T + tl) * p.E + ch;
                if (tl < p.T) {
                  p.dbg[o0] = (uint16_t)(gxr & 0xffffu); p.dbg[plane + o0] = (uint16_t)(gar & 0xffffu);
                  p.dbg[2 * plane + o0] = (uint16_t)(xc & 0xffffu);
                }
                if (tl + 1 < p.T) {
                  p.dbg[o0 + p.E] = (uint16_t)(gxr >> 16); p.dbg[plane + o0 + p.E] = (uint16_t)(gar >> 16);
                  p.dbg[2 * plane + o0 + p.E] = (uint16_t)(xc >> 16);
                }
              }
              A2[pi] = a2; X2[pi] = n2;
              const float al = bf_lo(a2), ah = bf_hi(a2);
              Hh = fmaf(al, Hh, bf_lo(n2));
              Hh = fmaf(ah, Hh, bf_hi(n2));
              P *= al; P *= ah;
            }
          }
          // this granule's TMEM slot may be overwritten by the next MMA
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(t_empty + slot);
          ++gq;
        }
        // ------------------------------------------------------ carry chain
        const size_t widx = (((size_t)fam * p.ntt + tt) * p.B + b) * kMch + chl;
        const size_t wstep = (size_t)p.B * kMch;           // one time tile back
        float c0;
        if (tt > 0 && tt + 1 < p.ntt) {   // tile 0 publishes its state right away instead
          st_relaxed_u64(p.agg_p + widx, pack_tagged(P, epoch));
          st_relaxed_u64(p.agg_h + widx, pack_tagged(Hh, epoch));
        }
        if (tt == 0) {
          c0 = p.h0 != nullptr ? p.h0[(size_t)b * p.E + ch] : 0.0f;
        } else {
          int j = tt - 1;
          size_t src = widx - wstep;
          unsigned long long pw;
          const long long t_start = clock64();
          unsigned polls = 0;
          for (;;) {
            pw = ld_relaxed_u64(p.pref + src);
            if (__all_sync(0xffffffffu, (unsigned)pw == epoch)) break;
            const bool agg = (unsigned)ld_relaxed_u64(p.agg_p + src) == epoch &&
                             (unsigned)ld_relaxed_u64(p.agg_h + src) == epoch;
            // (tile 0 never publishes an aggregate, only its state: j stays >= 0)
            if (__all_sync(0xffffffffu, agg)) { --j; src -= wstep; continue; }
            __nanosleep(20);
            polls += 63;   // a poll here costs ~a microsecond: check the watchdog every 16 polls
            if (__any_sync(0xffffffffu, watchdog_expired(t_start, p.err, 7, polls))) break;
          }
          c0 = tagged_value(pw);
          for (int k = j + 1; k < tt; ++k) {               // fold the aggregates passed on the way, left to right
            src += wstep;
            c0 = fmaf(tagged_value(ld_relaxed_u64(p.agg_p + src)), c0,
                      tagged_value(ld_relaxed_u64(p.agg_h + src)));
          }
        }
        if (tt + 1 < p.ntt) st_relaxed_u64(p.pref + widx, pack_tagged(fmaf(P, c0, Hh), epoch));
        // ------------------------------------------------------ pass 2 (replay)
        float h = c0;
        uint16_t* yp = p.y + ((size_t)b * p.T + t0) * p.E + ch;
        const bool full = t0 + TILE_T <= p.T;
#pragma unroll
        for (int pi = 0; pi < NP; ++pi) {
          const uint32_t a2 = A2[pi], n2 = X2[pi];
          float y0, y1;
          if constexpr (FAST) {
            y0 = fmaf(bf_lo(a2), h, bf_lo(n2));
            y1 = fmaf(bf_hi(a2), y0, bf_hi(n2));
          } else {                                         // mul then add, as the reference loop (:196)
            y0 = __fadd_rn(__fmul_rn(bf_lo(a2), h), bf_lo(n2));
            y1 = __fadd_rn(__fmul_rn(bf_hi(a2), y0), bf_hi(n2));
          }
          h = y1;
          const uint32_t o = pack_bf2(y0, y1);
          if (full || t0 + 2 * pi < p.T) st_u16(yp + (size_t)(2 * pi) * p.E, o);
          if (full || t0 + 2 * pi + 1 < p.T) st_u16(yp + (size_t)(2 * pi + 1) * p.E, o >> 16);
        }
        if (p.last_h != nullptr && tt == p.ntt - 1) p.last_h[(size_t)b * p.E + ch] = h;
      }
    }
  }

  // teardown: every role is done with TMEM
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace fused
}  // namespace cg
