// Fused RG-LRU for sm_100a: block-diagonal gate GEMMs on tcgen05 tensor cores
// with TMEM accumulators, gate math and the chunked linear-recurrence scan in
// the epilogue -- the gate pre-activations never reach HBM.
//
// Replaces, in ONE kernel, reference recurrentgemma/torch/layers.py:345-375
// (RGLRU.forward) INCLUDING the two BlockDiagonalLinear GEMMs (:133-142, :348-349)
// and rnn_scan (:146-199).  SURVEY.md section 8(f) row F1.
//
// Formulation (transposed GEMM, channels on TMEM lanes, time on TMEM columns):
//   work column = (batch row b, 128 channels of one head = a "family");  scan
//   tile = 32 time steps of one column.  One MMA tile covers TWO scan tiles of
//   different columns (N = 64: rows 0-31 and 32-63 of the X stage come from two
//   TMA boxes), so the 128 x K weight operand is fetched from shared memory
//   once per 64 steps.  Per MMA tile the MMA warp issues
//       D_x = Wx^T . X^T,   D_a = Wa^T . X^T   (M = 128 channels, K = head width)
//       D_t = I . X^T       (identity: the tensor core transposes x, exactly)
//   into 192 TMEM columns.  An epilogue thread owns ONE channel: tcgen05.ld
//   hands it the pre-activations and x of consecutive time steps in registers,
//   the gate math works on bf16x2 pairs (t, t+1), the per-step (a, x~) state is
//   parked in TMEM (tcgen05.st) between the two scan passes, and the recurrence
//   is a sequential fp32 scan in registers: no shared-memory staging of
//   activations at all.  Tiles of one column are chained through global memory
//   with the decoupled look-back of cg_scan.cuh (tagged 64-bit words, folds are
//   always left-to-right => bit-reproducible).
//
// Roles (640 threads, 1 CTA / SM, persistent):
//   warps 0-15   four epilogue warpgroups in two pairs; pair p owns TMEM columns
//                [256p, 256p+256): 192 accumulator + 2 x 32 state.  Warpgroup
//                (p, h) works on half h (columns 32h..32h+31) of every MMA tile
//                issued for pair p; MMA tiles alternate between the pairs.
//   warp 16      TMA producer and scheduler: packed, pre-swizzled gate weights once
//                per family, then two X boxes [32 steps x head width] per MMA tile
//                through a 3-D tensor map (SWIZZLE_128B, zero fill beyond T); writes
//                the tile descriptors (TileDesc) every other role follows; CONV
//                kernels: draws the tiles of long kernels from per-head ticket
//                counters and forwards the descriptors to the peer CTA of its cluster
//   warp 17      MMA issuer (one thread), tcgen05.commit -> mbarriers
//   warps 18-19  idle (registers are allocated per four warps).  CONV kernels: the
//                temporal convolution runs in the epilogue warpgroups, in place in
//                the X stage (conv_tile)
// CTA i works on family i % families; the CTAs of one family take its tiles
// two at a time, round-robin in time-major order, so look-back dependencies
// always point to tiles that are already running (all CTAs are co-resident:
// grid <= #SMs, 1 CTA / SM).  CONV kernels: the two CTAs that own the families of a
// head form a 2-CTA cluster (the convolution is computed once per head and the
// halves are exchanged); see Schedule (tail stealing) and the producer (tickets).
#pragma once

#include <cuda.h>

#include "cg_common.cuh"
#include "cg_scan.cuh"

// tuning switches (scripts/build_variants.py); the defaults are the shipped configuration
#ifndef CGF_SLEEP_NS
#define CGF_SLEEP_NS 64
#endif
// back-off of the producer / MMA / convolution warps' waits (their wake-up latency has slack)
#ifndef CGF_SLEEP_AUX_NS
#define CGF_SLEEP_AUX_NS CGF_SLEEP_NS
#endif
#ifndef CGF_PRELOAD
#define CGF_PRELOAD 1
#endif
// look-back depth per round trip (template parameter LOOK of the kernel; this is its default):
// 2 = the predecessor's state word plus the aggregates / state of the tile before it in ONE round trip.
// Measured (us, fused kernel alone, round 1): B=2,T=8192 141.6 -> 129.4; B=8,T=2048 107.0 -> 107.0;
// depth 4: 131.4 / 111.1 (more loads per poll than the chain is long).  Round 2, one-launch kernel at
// config 2 (B=8): depth 1 123.7 us vs depth 2 125.3 us -> the host picks 2 for B <= CGF_LOOK_SPLIT, else 1.
#ifndef CGF_LOOK
#define CGF_LOOK 2
#endif
#ifndef CGF_LOOK_SPLIT
#define CGF_LOOK_SPLIT 6    // batch sizes up to this use look-back depth 2 (B = 4: -4 %, B = 5, 6: neutral), larger ones depth 1
#endif
#ifndef CGF_EFIX
#define CGF_EFIX 1          // specialised instances with a compile-time row pitch for E = 2560
#endif
#ifndef CGF_KEEP_AGG
#define CGF_KEEP_AGG 1      // look-back: keep the accepted aggregates of the two nearest predecessors in registers
#endif
#ifndef CGF_HINT_NS
#define CGF_HINT_NS 20000
#endif
// CGF_ABLATE (timing experiments only, results are WRONG): 1 = no look-back,
// 2 = no y stores, 4 = no replay pass at all, 8 = no in-kernel convolution (loads, arithmetic, stores),
// 16 = no exchange of the convolved halves between the CTAs of a cluster
#ifndef CGF_ABLATE
#define CGF_ABLATE 0
#endif
// CGF_TRACE: debug_out becomes a timeline buffer [7 roles][1024] of
// (clock64 << 4 | event) words written by CTA 0 (scripts/fused_trace.py)
#ifndef CGF_TRACE
#define CGF_TRACE 0
#endif

#if CGF_TRACE
#define CGF_EVENT(role, code)                                                               \
  do {                                                                                      \
    if (blockIdx.x == 0 && lane == 0 && tn < 1024)                                          \
      reinterpret_cast<unsigned long long*>(p.dbg)[(role) * 1024 + tn++] =                  \
          (static_cast<unsigned long long>(clock64()) << 4) | (unsigned)(code);             \
  } while (0)
#else
#define CGF_EVENT(role, code) do { } while (0)
#endif

namespace cg {
namespace fused {

constexpr int kTile = 32;         // time steps per scan tile (one epilogue warpgroup)
constexpr int kMmaN = 2 * kTile;  // UMMA N: two scan tiles of different columns
constexpr int kMch = 128;         // channels per work column (UMMA M)
constexpr int kPairCols = 256;    // TMEM columns per warpgroup pair: 3 x 64 accumulator + 2 x 32 state
constexpr int kTmemCols = 512;
constexpr int kEpiWarps = 16;     // four warpgroups
constexpr int kAuxWarps = 4;      // fifth warpgroup: TMA producer, MMA issuer, two Conv1D warps
constexpr int kThreads = (kEpiWarps + kAuxWarps) * 32;
constexpr uint32_t kKBlockBytes = 128u * 128u;   // 128 rows x 64 bf16 (one swizzle-128B K block)

struct FusedParams {
  const unsigned char* wpack;      // [families][2 gates][KB][128 rows][128 B], swizzled (pack kernel)
  const unsigned char* ident;      // [2][128 rows][128 B] identity, swizzled
  const uint16_t* bias_x;          // [E] or null
  const uint16_t* bias_a;
  const uint16_t* neg8sp_bf;       // [E] -8*softplus(a_param) as bf16 (prologue)
  const unsigned* reset_bits;      // [rows][words] bit t%32 of word t/32
  long long bits_bstride;          // words per batch row (0 = broadcast)
  int words;                       // words per row (= ntt: one word per 32-step tile)
  const float* h0;                 // [B,E] or null
  uint16_t* y;                     // [B,T,E]
  float* last_h;                   // [B,E] or null
  const unsigned* epoch;
  unsigned long long* agg_p;       // [families][ntt][B][128]
  unsigned long long* agg_h;
  unsigned long long* pref;
  // in-kernel temporal convolution (CONV kernels): the tensor map then describes x_lin, the INPUT of
  // Conv1D.forward, and the X stages are convolved in place before the MMAs read them
  const uint16_t* x_lin;           // [B,T,E] (halo rows and the returned cache are read directly)
  const uint16_t* conv_w;          // [4,E]
  const uint16_t* conv_b;          // [E]
  uint16_t* conv_cache;            // optional [B,3,E]: last three input rows, left zero padded (layers.py:542-543)
  int mask_mode;                   // CG_MASK_FORK / CG_MASK_UPSTREAM
  const uint16_t* gate_mul;        // optional [B,T,E]: y <- round_bf16(y * gate_mul) (RecurrentBlock, modules.py:651)
  uint16_t* dbg;                   // optional [3][B][T][E]: rounded pre_x, pre_a, x^T
  int* err;                        // watchdog flag
  int B, T, E;
  int ntt;                         // time tiles per batch row
  int families;                    // E / 128
  // ticket / B without the ~25-instruction emulated division (Granlund-Montgomery, exact for every
  // 32-bit ticket): q = (t + ((n - t) >> s1)) >> s2 with t = umulhi(m, n)   (make_bdiv on the host)
  uint32_t bdiv_m, bdiv_s1, bdiv_s2;
  int steal;                       // pairs a poor family hands to helper CTAs (Schedule), 0 = none
  int* tickets;                    // CONV kernels: one ticket counter per head, zeroed by the prologue kernel
  int forward;                     // CONV clusters: the leader CTA's producer forwards the tile descriptors to its peer
                                   // (required for static_pct != 0; a static schedule may also be derived by both CTAs)
  int static_pct;                  // CONV kernels: 0 = static schedule (Schedule, with `steal`); 1..100 = that share of a
                                   // head's tiles round-robin among its home clusters, the rest through the tickets
};
__host__ __device__ __forceinline__ void make_bdiv(uint32_t d, uint32_t& m, uint32_t& s1, uint32_t& s2) {
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;                               // ceil(log2 d)
  m = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
  s1 = l < 1 ? l : 1;
  s2 = l < 1 ? 0 : l - 1;
}

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
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"((unsigned)CGF_HINT_NS) : "memory");   // suspend-time hint (ns)
  return ok != 0;
}
// Bounded wait: a protocol bug must end the launch, never hang the device, and
// must never go unnoticed.  On a timeout the wait records which wait it was in
// the workspace header (diagnosis under a debugger) and TRAPS: the launch fails
// and the stream carries a sticky CUDA error that the host sees at its next
// synchronisation -- no call can return garbage with status 0.
constexpr long long kWatchdogCycles = 4000000000LL;   // ~2 s
__device__ __forceinline__ bool watchdog_expired(long long t0, int* err, int code, unsigned& polls) {
  if ((++polls & 63u) != 0u) return false;
  if (clock64() - t0 > kWatchdogCycles) {
    atomicCAS(err, 0, code);
    __threadfence_system();
    __trap();
  }
  return false;
}
// NS: plain back-off between two polls.  The loop body is kept minimal (a waiting warp's
// polls compete for issue slots with the warps of its scheduler that have work) and the
// watchdog's clock is read once per 64 polls only.
template <int NS = CGF_SLEEP_NS>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err, int code) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
      if (NS > 0) __nanosleep(NS);
      if (mbar_try_wait(bar, parity)) return;
    }
    if (clock64() - t0 > kWatchdogCycles) {
      atomicCAS(err, 0, code);
      __threadfence_system();
      __trap();
    }
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
// One lane of a CONVERGED warp.  The roles that issue TMA / tcgen05.mma run with
// all 32 lanes so that descriptors and coordinates stay warp-uniform (uniform
// registers); only the issuing instruction itself is predicated on the elected
// lane.  (Inside an `if (lane == 0)` region the compiler cannot prove uniformity
// and wraps every UTCHMMA in a R2UR broadcast loop -- ~100 cycles per MMA.)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}"
               : "=r"(pred) :: "memory");
  return pred != 0;
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
// The same MMA with the descriptors given as their LOW words (start address >> 4 | LBO): the high
// word (SBO = 1024 B, version 1, SWIZZLE_128B) is the same constant for every operand of this
// kernel, and advancing an operand by `bytes` is ONE 32-bit add of bytes >> 4 to the low word
// (all of shared memory is below 256 KB, so the 14-bit address field cannot overflow).  Building
// each 64-bit descriptor from the address (shift, mask, or, or) made the issue loop spend ~10
// uniform-datapath instructions per MMA -- as long as the MMA itself takes.
constexpr uint32_t kDescHi = 0x40004040u;   // (64 << 0) | (1 << 14) | (2 << 29)  ==  bits 32.. of umma_desc()
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr) {
  return ((smem_addr >> 4) & 0x3fffu) | (1u << 16);
}
__device__ __forceinline__ void umma_bf16_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
      :: "r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               :: "r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 8 consecutive 32-bit columns: thread i of the warp owns lane
// (quadrant base + i), registers = columns.
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// fp32 + bf16 half of a packed register in one instruction (FHADD.BF16 with a
// half selector on sm_100a): no unpack, one rounding -- == fadd_rn(widen(x), c).
__device__ __forceinline__ float add_bf_lo(uint32_t packed, float c) {
  float d;
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tadd.rn.f32.bf16 %0, lo, %2;\n\t}"
      : "=f"(d) : "r"(packed), "f"(c));
  return d;
}
__device__ __forceinline__ float add_bf_hi(uint32_t packed, float c) {
  float d;
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tadd.rn.f32.bf16 %0, hi, %2;\n\t}"
      : "=f"(d) : "r"(packed), "f"(c));
  return d;
}
__device__ __forceinline__ void tmem_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t ld_u16(const uint16_t* p) {
  uint16_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_u16(uint16_t* p, uint32_t v) {
  // (cache policies .cs / .wt / L2::evict_first were measured: no change of the step time)
  asm volatile("st.global.u16 [%0], %1;" :: "l"(p), "h"(static_cast<uint16_t>(v)) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 r;   // volatile, no "memory" clobber: ordered against sts128, free against arithmetic
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- thread-block cluster helpers (CONV kernels at head width 256: the two CTAs that own the two
// channel halves of a head form a cluster and exchange their halves of the convolved X stage)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// bulk copy local shared memory -> a peer CTA's shared memory, completing (bytes) on the PEER's mbarrier
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes,
                                                  uint32_t bar_cluster) {
  asm volatile(
      "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      :: "r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster) : "memory");
}
// tcgen05.commit that arrives on the barrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               :: "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t r;   // volatile, no "memory" clobber: ordered against sts32, free against arithmetic
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr));
  return r;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v));
}
__device__ __forceinline__ uint32_t ldg32_nc(const void* p) {
  uint32_t r;
  asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
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

// Per-tile bookkeeping, computed ONCE per CTA by the (otherwise idle) TMA producer warp and handed to
// the sixteen epilogue warps through shared memory: scan tile (tt, b) of a ticket, the tile's document
// starts, its element offset in the [B,T,E] tensors and its exchange-word index.  The epilogue warps
// used to derive all of this themselves, every warp for every tile: ~200 of a warp's ~1 600 instructions
// per tile at an IPC of ~0.1 (serial integer chains).  The descriptor of pair p's MMA tile with use
// count u sits in slot [p][u & 3][half].  It is written ONE TILE AHEAD: before the producer's arrive
// (release) on raw_full[p] for use u - 1 (u = 0: for use 0), so a warpgroup that has waited on
// raw_full[p] for use u - 1 (acquire) may read it -- early enough to request the halo rows of the next
// convolution a whole tile before they are needed.  A slot is rewritten four uses later.
struct TileDesc {
  int tt;                 // time tile, -1: this half of the MMA tile is empty (odd tile count), -2: STOP (no more tiles)
  int b;                  // batch row
  uint32_t rbits;         // document starts inside the tile (bit t % 32)
  uint32_t rprev;         // ... of the previous tile of the row (0 before the first)
  uint32_t eoff_lo, eoff_hi;   // element offset of (b, tt * 32, channel 0) in x / y
  uint32_t widx;          // exchange-word index of (family, tt, b, channel 0 of the family)
  int fam;                // family (the dynamic schedule of the CONV kernels hands out tiles of several)
};
static_assert(sizeof(TileDesc) == 32, "two 16-byte shared-memory accesses");

template <int KB>
struct FusedCfg {
  static constexpr uint32_t kWBytes = 2u * KB * kKBlockBytes;
  static constexpr uint32_t kIBytes = 2u * kKBlockBytes;
  static constexpr uint32_t kXKBlock = kMmaN * 128u;               // one K block of an X stage: 64 rows x 128 B
  static constexpr uint32_t kXStageBytes = static_cast<uint32_t>(KB) * kXKBlock;
  static constexpr int kXStages = 2;                               // one per warpgroup pair
  static constexpr int kBars = 2 + 3 * kXStages + 4;
  static constexpr uint32_t kTapBytes = 5u * 128u * 2u;             // CONV: w[0..3], b of this CTA's 128 input channels
  static constexpr uint32_t kDescBytes = 2u * 4u * 2u * 32u;        // tile descriptors [pair][use & 3][half], see TileDesc
  static constexpr size_t kSmemBytes = 1024 + kWBytes + kIBytes + kXStages * kXStageBytes + kBars * 8 + 16 + kTapBytes +
                                       kDescBytes + 32;   // + dsc_full[2][2] (CONV clusters)
  static_assert(KB % 2 == 0, "head width must be a multiple of 128");
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
// Work schedule, identical in every role of a CTA.  The ticket pairs (= MMA
// tiles) of a family are numbered j = 0 .. npairs-1 in time-major order; a CTA
// works through SEGMENTS (family, first pair, stride, count), in order, and
// reloads the gate weights when the family changes.
//   G < families           CTA i takes families i, i+G, ... completely (test hook)
//   G = d * families + r   (148 SMs, 20 families: d = 7, r = 8): the d or d + 1 CTAs of
//       a family share its pairs round-robin, so every look-back dependency points
//       to a pair that is running or done.
// (A balanced variant with "floater" CTAs that help several families was measured
// slower -- 116 us vs 105.6 us at config 2: a floater couples the chains of the
// families it visits -- and removed)
// ---------------------------------------------------------------------------
struct Seg { int fam, j0, stride, count; };
// Round-robin mode with TAIL STEALING (steal > 0).  G = d * nfam + r: families f < r ("rich") have d + 1
// CTAs, the others ("poor") d -- at config 2 (148 CTAs, 20 families; one launch: 74 clusters, 10 heads)
// 32 vs 37 MMA tiles per CTA, and the poor ones set the kernel's time.  Each poor family therefore hands
// its LAST `steal` pairs to CTAs of rich families: rich CTA number rho (of nrich) helps poor family
// r + rho % npoor as helper rho / npoor and takes every nh-th of those pairs, after its own work and a
// weight reload.  The stolen tiles are the end of the poor family's time line: nobody but their own
// successors waits for them, so a late helper delays only that tail (the static "floater" of round 1, which
// joined a family in the MIDDLE of its time line, stalled seven other CTAs whenever it was late).  The
// host picks `steal` so that both kinds of CTA end at the same time (fused_forward).
struct Schedule {
  int mode;        // 0 = whole families, 1 = round-robin
  int cta, G, nfam, npairs, steal;
  __host__ __device__ __forceinline__ Schedule(int cta_, int G_, int nfam_, int npairs_, int steal_ = 0)
      : cta(cta_), G(G_), nfam(nfam_), npairs(npairs_), steal(steal_) {
    mode = G < nfam ? 0 : 1;
    const int r = G % nfam, d = G / nfam;
    if (mode == 0 || r == 0 || steal < 0 || steal > npairs || r * (d + 1) < nfam - r) steal = 0;
  }
  __host__ __device__ __forceinline__ bool rich() const { return cta % nfam < G % nfam; }
  __host__ __device__ __forceinline__ int nseg() const {
    if (mode == 0) return cta < nfam ? (nfam - 1 - cta) / G + 1 : 0;
    return steal > 0 && rich() ? 2 : 1;
  }
  __host__ __device__ __forceinline__ Seg get(int s) const {
    Seg sg;
    if (mode == 0) {
      sg.fam = cta + s * G; sg.j0 = 0; sg.stride = 1; sg.count = npairs;
    } else if (s == 0) {
      sg.fam = cta % nfam;
      const int rank = cta / nfam;
      const int own = rich() ? npairs : npairs - steal;    // a poor family's own CTAs stop before its tail
      sg.stride = (G - 1 - sg.fam) / nfam + 1;
      sg.j0 = rank;
      sg.count = rank < own ? (own - rank + sg.stride - 1) / sg.stride : 0;
    } else {
      const int r = G % nfam, d = G / nfam, npoor = nfam - r, nrich = r * (d + 1);
      const int rho = (cta / nfam) * r + cta % nfam;       // 0 .. nrich - 1
      const int pf = rho % npoor, hi = rho / npoor;        // poor family r + pf, helper number hi
      const int nh = (nrich - pf + npoor - 1) / npoor;     // helpers of that family
      sg.fam = r + pf;
      sg.stride = nh;
      sg.j0 = npairs - steal + hi;
      sg.count = hi < steal ? (steal - hi + nh - 1) / nh : 0;
    }
    return sg;
  }
};

// ---------------------------------------------------------------------------
// In-kernel temporal convolution (CONV kernels; north star: "the width-4 Conv1D is
// fused into the same pass").  The TMA producer loads the rows of x_lin -- the
// INPUT of Conv1D.forward -- into the X stage and the EPILOGUE warpgroups convolve
// them in place before the MMA warp may read the stage (conv_tile in the kernel):
// the convolution's issue load is spread over the SM's four sub-partitions, and at
// head width 256 the two CTAs of a head (a 2-CTA cluster) each convolve their 128
// channels and exchange the halves with bulk shared -> shared::cluster copies, so
// the convolution is computed once per head.  Arithmetic: the reference's
// accumulation order with one bf16 rounding per eager op (layers.py:530-536:
// packed HMUL2 / HADD2, no contraction), bit-exact with cg::conv1d_w4_kernel and
// the reference.  The conv output never reaches HBM.
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// The fused kernel.  KB = head width / 64 (K blocks of the gate GEMMs).
// ---------------------------------------------------------------------------
// 20 warps are allocated with 96 registers each; setmaxnreg then moves registers
// from the producer / MMA / convolution warpgroup to the four epilogue warpgroups.
// MUL: multiply the output by a second activation tensor on the way out (the
// gating product of RecurrentBlock), SURVEY.md section 8(f) row F2.
// CONV: the temporal convolution runs in the kernel (see above): `tmap_x` describes
// x_lin and the kernel computes Conv1D.forward -> RGLRU.forward in one launch.
// LOOK: look-back depth per round trip (see resolve_carry): 2 for small batches, where the carry chain
// is the bound; 1 where the chain has slack (fewer loads per poll).
// EFIX: the row pitch E as a compile-time constant (0 = p.E): the y stores of the replay pass and the halo
// loads of the convolution then use immediate offsets instead of 64-bit address arithmetic per row.
// UNI: the warp index goes through a shuffle, which tells the compiler that the role branches are
// warp-uniform (see below): 2.6 instructions per element fewer, 2-7 % faster for dynamic schedules and small
// batches, 1.5 % slower for static schedules at B >= 8 (profiles/r3_ab_uniform_warp_index.txt) -- the host picks.
template <int KB, bool FAST, bool DBG, bool MUL, bool CONV, int LOOK = CGF_LOOK, int EFIX = 0, bool UNI = false>
__global__ void __launch_bounds__(kThreads, 1)
rglru_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const FusedParams p) {
  const int rowE = EFIX > 0 ? EFIX : p.E;
  using Cfg = FusedCfg<KB>;
  constexpr int CBS = KB / 2;               // 128-channel halves (families) per head
  constexpr int XS = Cfg::kXStages;
  constexpr uint32_t IDESC = umma_idesc(kMch, kMmaN);
  // CONV: the CBS CTAs that own the families of one head form a thread-block cluster.  Each CTA loads
  // and convolves only ITS 128 input channels (KBL = 2 K blocks of the X stage) and sends the result
  // to its peer with one bulk shared->shared copy per K block, so the convolution is computed once per
  // head, not once per family.
  constexpr int CL = CONV ? CBS : 1;        // cluster size
  constexpr int KBL = KB / CL;              // K blocks of an X stage this CTA loads (and convolves)
  static_assert(!CONV || KBL == 2, "the in-kernel convolution maps one warp to one K block of a half stage");

  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  unsigned char* sm = smem_raw + (base - raw_addr);
  const uint32_t sW = base, sI = sW + Cfg::kWBytes, sX = sI + Cfg::kIBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + Cfg::kWBytes + Cfg::kIBytes + XS * Cfg::kXStageBytes);
  uint64_t* w_full = bars;
  uint64_t* w_empty = bars + 1;
  uint64_t* raw_full = bars + 2;            // [XS] the TMA boxes of the stage have landed
  uint64_t* x_full = raw_full + XS;         // [XS] CONV: both halves of the stage are convolved
  uint64_t* x_empty = x_full + XS;          // [XS] the MMAs that read the stage are complete
  uint64_t* t_full = x_empty + XS;          // [2] accumulators of pair p are complete
  uint64_t* t_empty = t_full + 2;           // [2] both warpgroups of pair p have read them
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(t_empty + 2);
  const uint32_t sTap = smem_u32(tmem_holder) + 16u;   // [5][128] bf16: conv taps w[0..3], bias (CONV, one family per CTA)
  const uint32_t sDesc = sTap + Cfg::kTapBytes;        // TileDesc [pair][use & 3][half]
  // dsc_full[pair][use & 1] (CONV clusters): the leader CTA's producer has written the descriptors of
  // pair `pair`'s tile number `use` into THIS CTA's ring (see the producer)
  uint64_t* dsc_full = reinterpret_cast<uint64_t*>(sm + Cfg::kWBytes + Cfg::kIBytes + XS * Cfg::kXStageBytes +
                                                   Cfg::kBars * 8 + 16 + Cfg::kTapBytes + Cfg::kDescBytes);
  auto desc_addr = [&](uint32_t pair, uint32_t use, uint32_t half) -> uint32_t {
    return sDesc + (((pair * 4u + (use & 3u)) * 2u + half) << 5);
  };
  uint64_t* mma_ready = CONV ? x_full : raw_full;   // what the MMA warp waits for

  // (UNI: through a shuffle -- the compiler then knows that the role branches below are warp-uniform and may
  // keep uniform registers, e.g. the global-memory descriptor of every LDG / STG, live across them; with
  // threadIdx.x >> 5 it re-materialises that descriptor with two R2UR before each access: 2.6 instructions
  // per element)
  const int warp = UNI ? __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0) : (int)(threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;

#if CGF_TRACE
  if (threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    reinterpret_cast<unsigned long long*>(p.dbg)[7 * 1024 + 256 + blockIdx.x] = t;
  }
#endif
  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    mbar_init(w_empty, 1);
    // x_full: one arrival per warp of the pair's two warpgroups (+ the peer's bytes, CL == 2); x_empty: the MMAs of
    // EVERY CTA of the cluster have read the stage (the peer writes into my stage too)
    for (int i = 0; i < XS; ++i) { mbar_init(raw_full + i, 1); mbar_init(x_full + i, 8); mbar_init(x_empty + i, CL); }
    for (int i = 0; i < 2; ++i) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, 8); }
    for (int i = 0; i < 4; ++i) mbar_init(dsc_full + i, 1);
    fence_mbar_init();
  }
  if (warp == kEpiWarps) tmem_alloc(smem_u32(tmem_holder), kTmemCols);
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();   // the peer's barriers exist before anything is sent to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;


  // work schedule of this CTA (identical in every role), see Schedule above.
  // Ticket = tt * B + b (time-major); pair j = tickets 2j, 2j + 1 = one MMA tile.
  const int nfam = p.families;
  const int ntiles = p.ntt * p.B;
  const int npairs = (ntiles + 1) >> 1;
  // CONV: the schedule hands out HEADS to clusters; CTA `crank` of the cluster takes family head * CBS + crank
  const Schedule sched(CONV ? (int)blockIdx.x / CL : (int)blockIdx.x, CONV ? (int)gridDim.x / CL : (int)gridDim.x,
                       CONV ? nfam / CBS : nfam, npairs, p.steal);
  const int nsegs = sched.nseg();
  auto seg_family = [&](const Seg& sg) -> int { return CONV ? sg.fam * CBS + (int)crank : sg.fam; };
  auto div_b = [&](int n) -> int {                           // n / p.B for 0 <= n
    const uint32_t t = __umulhi(p.bdiv_m, (uint32_t)n);
    return (int)((t + (((uint32_t)n - t) >> p.bdiv_s1)) >> p.bdiv_s2);
  };

  if (warp == kEpiWarps) {
    // ===================================================== TMA producer
    if constexpr (!CONV) {
      uint32_t mq = 0, witer = 0;
      int tn = 0; (void)tn;
      int cur_fam = -1;
      for (int sgi = 0; sgi < nsegs; ++sgi) {
        const Seg sg = sched.get(sgi);
        if (sg.count == 0) continue;
        const int fam = seg_family(sg);
        if (fam != cur_fam) {
        cur_fam = fam;
        if (witer > 0) mbar_wait<CGF_SLEEP_AUX_NS>(w_empty, (witer - 1) & 1, p.err, 1);
        const unsigned char* wsrc = p.wpack + (size_t)fam * Cfg::kWBytes;
        if (elect_one()) {
          mbar_expect_tx(w_full, Cfg::kWBytes + Cfg::kIBytes);
#pragma unroll 1
          for (uint32_t off = 0; off < Cfg::kWBytes; off += kKBlockBytes)
            bulk_load(sW + off, wsrc + off, kKBlockBytes, w_full);
#pragma unroll 1
          for (uint32_t off = 0; off < Cfg::kIBytes; off += kKBlockBytes)
            bulk_load(sI + off, p.ident + off, kKBlockBytes, w_full);
        }
        __syncwarp();
        // x is written by the kernel(s) before the prologue in stream order; as a
        // programmatic dependent this grid may be running before they are done.
        // The weights above do not depend on them, the X boxes below do.
        if (witer == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
        ++witer;
        }
        const int c_head = (fam / CBS) * (KB * 64);
#pragma unroll 1
        for (int m = 0; m < sg.count; ++m, ++mq) {
          const int t1st = 2 * (sg.j0 + m * sg.stride);
          const int nhalf = t1st + 1 < ntiles ? 2 : 1;
          const uint32_t stage = mq & 1u, use = mq >> 1;
          // Tile descriptors (TileDesc): lanes 0 / 1 = the halves of THIS tile (first use of the pair
          // only), lanes 2 / 3 = the halves of the pair's NEXT tile (two MMA tiles ahead); the reset
          // words are requested before the wait below.
          TileDesc td{-1, 0, 0u, 0u, 0u, 0u, 0u, 0};
          if (lane < 4) {
            int dfam = fam, dj = sg.j0 + m * sg.stride;
            bool have = lane < 2;
            if (lane >= 2) {                               // the tile two positions after (sgi, m)
              int s2 = sgi, m2 = m + 2;
              Seg g2 = sg;
              have = true;
              while (m2 >= g2.count) {
                m2 -= g2.count;
                do {
                  ++s2;
                  if (s2 < nsegs) g2 = sched.get(s2);
                } while (s2 < nsegs && g2.count == 0);
                if (s2 >= nsegs) { have = false; break; }
              }
              dfam = seg_family(g2); dj = g2.j0 + m2 * g2.stride;
            }
            const int ticket = 2 * dj + (lane & 1);
            if (have && ticket < ntiles) {
              td.tt = div_b(ticket); td.b = ticket - td.tt * p.B;
              const unsigned* rw = p.reset_bits + (long long)td.b * p.bits_bstride + td.tt;
              td.rbits = rw[0];
              td.rprev = td.tt > 0 ? rw[-1] : 0u;
              const unsigned long long eo = ((unsigned long long)td.b * p.T + (unsigned long long)td.tt * kTile) * p.E;
              td.eoff_lo = (uint32_t)eo; td.eoff_hi = (uint32_t)(eo >> 32);
              td.widx = (uint32_t)((((size_t)dfam * p.ntt + td.tt) * p.B + td.b) * kMch);
            }
            td.fam = dfam;
          }
          CGF_EVENT(0, 1);
          mbar_wait<CGF_SLEEP_AUX_NS>(x_empty + stage, (use & 1) ^ 1, p.err, 2);
          CGF_EVENT(0, 2);
          if (lane < 4 && (lane >= 2 || use == 0)) {
            const uint32_t da = desc_addr(stage, use + (uint32_t)(lane >> 1), (uint32_t)(lane & 1));
            sts128(da, make_uint4((uint32_t)td.tt, (uint32_t)td.b, td.rbits, td.rprev));
            sts128(da + 16, make_uint4(td.eoff_lo, td.eoff_hi, td.widx, (uint32_t)td.fam));
          }
          __syncwarp();
          if (elect_one()) mbar_expect_tx(raw_full + stage, nhalf * (Cfg::kXStageBytes / 2 / CL));
          for (int hf = 0; hf < nhalf; ++hf) {
            const int ticket = t1st + hf;
            const int tt = div_b(ticket), b = ticket - tt * p.B;
            if (elect_one()) {
#pragma unroll
              for (int kbl = 0; kbl < KBL; ++kbl) {
                const int kb = (CL > 1 ? (int)crank * KBL : 0) + kbl;   // CONV: my channel half only
                tma_load_3d(sX + stage * Cfg::kXStageBytes + kb * Cfg::kXKBlock + hf * (kTile * 128), &tmap_x,
                            raw_full + stage, c_head + kb * 64, tt * kTile, b);
              }
            }
          }
          __syncwarp();
        }
      }
    } else {
      // ---------------------------------------------------------------------------------------------
      // CONV kernels: DYNAMIC schedule.  The tiles of a head are handed out in time-major order by a
      // ticket counter per head in global memory (zeroed by the prologue kernel): a cluster draws from its
      // home head until that is exhausted and then from the other heads, so no cluster idles while tiles
      // are left -- the static round-robin leaves the clusters of the heads that own one cluster more
      // (74 clusters, 10 heads) idle for the last 11 % of the kernel.  A ticket's predecessors in time were
      // drawn earlier, i.e. are running or done, whoever holds them: the look-back cannot deadlock.  Only
      // the LEADER CTA of a cluster draws (lane 0 of this warp); it writes the tile descriptors into its own
      // ring AND into the peer's (st.shared::cluster), two tiles ahead, and arrives on the peer's dsc_full;
      // every other role of both CTAs just follows the descriptors, a descriptor with tt == -2 ends them.
      // ---------------------------------------------------------------------------------------------
      const int nheads = nfam / CBS;
      const int cluster = (int)blockIdx.x / CL, nclusters = (int)gridDim.x / CL;
      const int home = cluster % nheads;
      // static schedule: both CTAs of a cluster derive the same tile sequence on their own (nothing to
      // forward); dynamic tickets: the leader CTA draws and forwards the descriptors to its peer
      const bool forward = CL > 1 && p.forward != 0;
      const bool leader = crank == 0 || !forward;
      uint32_t witer = 0;
      int tn = 0; (void)tn;
      int cur_head = -1;
      auto load_weights = [&](int head) {
        cur_head = head;
        if (witer > 0) mbar_wait<CGF_SLEEP_AUX_NS>(w_empty, (witer - 1) & 1, p.err, 1);
        const unsigned char* wsrc = p.wpack + (size_t)(head * CBS + (int)crank) * Cfg::kWBytes;
        if (elect_one()) {
          mbar_expect_tx(w_full, Cfg::kWBytes + Cfg::kIBytes);
#pragma unroll 1
          for (uint32_t off = 0; off < Cfg::kWBytes; off += kKBlockBytes)
            bulk_load(sW + off, wsrc + off, kKBlockBytes, w_full);
#pragma unroll 1
          for (uint32_t off = 0; off < Cfg::kIBytes; off += kKBlockBytes)
            bulk_load(sI + off, p.ident + off, kKBlockBytes, w_full);
        }
        __syncwarp();
        // x_lin and the prologue's outputs (ticket counters, reset bitmask): see the static producer
        if (witer == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
        ++witer;
      };
      load_weights(home);                                  // the first tiles are the home head's
      // ---- tickets (leader): (head, pair index) or head < 0 = STOP
      // HYBRID: the first CGF_STATIC_PCT % of a head's pairs are dealt round-robin among its home clusters
      // (clusters in lock-step: a tile's predecessors are processed half a round earlier, the look-back
      // rarely waits); only the rest goes through the counter -- purely dynamic tickets, drawn two tiles
      // ahead, scramble the order in which the tiles of a chain are processed (133 vs 122 us at config 2).
      auto n_home = [&](int head) -> int { return head < nclusters ? (nclusters - 1 - head) / nheads + 1 : 0; };
      auto j_static = [&](int head) -> int {
        const int nh = n_home(head);
        return nh > 0 ? (int)((long long)npairs * p.static_pct / 100 / nh) * nh : 0;
      };
      const int my_rank = cluster / nheads, my_nh = n_home(home), my_js = j_static(home);
      int own_k = 0;
      int fails = 0;
      bool dry = false;
      int s_sgi = 0, s_m = 0;                              // p.static_pct == 0: cursor over the static schedule
      Seg s_sg = nsegs > 0 ? sched.get(0) : Seg{0, 0, 1, 0};
      auto draw = [&](int& head_out, int& j_out) {
        head_out = -1; j_out = 0;
        if (p.static_pct == 0) {
          // the STATIC schedule (round-robin + tail stealing, cg::fused::Schedule) through the same machinery:
          // the better choice for short kernels (config 2: 121.8 us vs 126.0 us with 10 % dynamic tickets)
          while (s_sgi < nsegs) {
            if (s_m < s_sg.count) { head_out = s_sg.fam; j_out = s_sg.j0 + s_m * s_sg.stride; ++s_m; return; }
            ++s_sgi; s_m = 0;
            if (s_sgi < nsegs) s_sg = sched.get(s_sgi);
          }
          return;
        }
        if (own_k >= 0) {                                  // static share of the home head
          const int j = my_rank + own_k * my_nh;
          if (j < my_js) { ++own_k; head_out = home; j_out = j; return; }
          own_k = -1;
        }
        while (!dry) {
          // which heads still have tickets: one relaxed load per lane (a failed atomicAdd per exhausted head
          // costs a round trip each -- ~10 us at the end of the kernel when every cluster runs dry at once)
          int left = 0;
          if (lane < nheads) {
            int taken;
            asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(taken) : "l"(p.tickets + lane) : "memory");
            left = npairs - j_static(lane) - taken;
          }
          const unsigned avail = __ballot_sync(0xffffffffu, left > 0);
          if (avail == 0u) { dry = true; break; }
          // the home head first, then the others starting at a cluster-specific offset (the clusters of
          // one head fan out over the other heads)
          int pick = home;
          if (!((avail >> home) & 1u)) {
            const int start = (home + 1 + my_rank) % nheads;
            const unsigned rot = ((avail >> start) | (avail << (nheads - start))) & (nheads < 32 ? (1u << nheads) - 1u : ~0u);
            pick = (start + (__ffs(rot) - 1)) % nheads;
          }
          int j = 0;
          if (lane == 0) j = atomicAdd(p.tickets + pick, 1);
          j = __shfl_sync(0xffffffffu, j, 0) + j_static(pick);
          if (j < npairs) { head_out = pick; j_out = j; return; }
          if (++fails > 4 * nheads) { dry = true; break; }   // (cannot happen: every failure exhausts a head for good)
        }
      };
      int h0 = -1, j0 = 0, h1 = -1, j1 = 0;              // tickets of tiles mq, mq + 1
      if (leader) { draw(h0, j0); draw(h1, j1); }
      uint32_t nstop = 0;
      const uint32_t peer = crank ^ 1u;
#pragma unroll 1
      for (uint32_t mq = 0;; ++mq) {
        const uint32_t stage = mq & 1u, use = mq >> 1;
        if (leader) {
          int h2, j2;
          draw(h2, j2);                                    // tile mq + 2
          // descriptors: lanes 0 / 1 = the halves of THIS tile (first use of a pair only), lanes 2 / 3 =
          // the halves of tile mq + 2; the reset words are requested before the wait below
          TileDesc td{-2, 0, 0u, 0u, 0u, 0u, 0u, 0};
          if (lane < 4 && (lane >= 2 || use == 0)) {
            const int dh = lane < 2 ? h0 : h2, dj = lane < 2 ? j0 : j2;
            if (dh >= 0) {
              td.tt = -1;
              td.fam = dh * CBS + (int)crank;              // my family; a forwarded copy gets the peer's
              const int ticket = 2 * dj + (lane & 1);
              if (ticket < ntiles) {
                td.tt = div_b(ticket); td.b = ticket - td.tt * p.B;
                const unsigned* rw = p.reset_bits + (long long)td.b * p.bits_bstride + td.tt;
                td.rbits = rw[0];
                td.rprev = td.tt > 0 ? rw[-1] : 0u;
                const unsigned long long eo = ((unsigned long long)td.b * p.T + (unsigned long long)td.tt * kTile) * p.E;
                td.eoff_lo = (uint32_t)eo; td.eoff_hi = (uint32_t)(eo >> 32);
                td.widx = (uint32_t)((((size_t)td.fam * p.ntt + td.tt) * p.B + td.b) * kMch);
              }
            }
          }
          // The ring slot of tile use + 1 held tile use - 3: every reader of that one is done once the MMAs
          // of use - 2 are complete in both CTAs, which this warp saw two iterations ago -- so the
          // descriptors (and the signal to the peer) go out BEFORE the wait for the stage, off the critical path
          if (lane < 4 && (lane >= 2 || use == 0)) {
            const uint32_t da = desc_addr(stage, use + (uint32_t)(lane >> 1), (uint32_t)(lane & 1));
            sts128(da, make_uint4((uint32_t)td.tt, (uint32_t)td.b, td.rbits, td.rprev));
            sts128(da + 16, make_uint4(td.eoff_lo, td.eoff_hi, td.widx, (uint32_t)td.fam));
            if (forward) {                                 // the peer's copy: its family is the next one
              const uint32_t dr = mapa_u32(da, peer);
              const uint32_t wpeer = td.widx + (uint32_t)p.ntt * (uint32_t)p.B * kMch;
              asm volatile("st.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(dr), "r"((uint32_t)td.tt),
                           "r"((uint32_t)td.b), "r"(td.rbits), "r"(td.rprev) : "memory");
              asm volatile("st.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(dr + 16), "r"(td.eoff_lo),
                           "r"(td.eoff_hi), "r"(wpeer), "r"((uint32_t)(td.fam + 1)) : "memory");
            }
          }
          __syncwarp();
          if (forward) {
            if (lane == 0) {                               // the peer's producer may read them now
              asm volatile("fence.acq_rel.cluster;" ::: "memory");
              if (use == 0)
                asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];"
                             :: "r"(mapa_u32(smem_u32(dsc_full + stage * 2), peer)) : "memory");
              asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];"
                           :: "r"(mapa_u32(smem_u32(dsc_full + stage * 2 + ((use + 1) & 1)), peer)) : "memory");
            }
            __syncwarp();
          }
          h0 = h1; j0 = j1; h1 = h2; j1 = j2;
          CGF_EVENT(0, 1);
          mbar_wait<CGF_SLEEP_AUX_NS>(x_empty + stage, (use & 1) ^ 1, p.err, 2);
          CGF_EVENT(0, 2);
        } else {
          // the peer: the descriptors arrive from the leader -- this tile's (first use of a pair) and the
          // look-ahead tile's (the epilogue warpgroups read it as soon as raw_full of THIS tile completes)
          auto wait_desc = [&](uint32_t u) {
            uint64_t* bar = dsc_full + stage * 2 + (u & 1);
            const uint32_t parity = (u >> 1) & 1;
            uint32_t ok = 0;
            const long long t0c = clock64();
            unsigned polls = 0;
            for (;;) {
              asm volatile("{\n\t.reg .pred p;\n\t"
                           "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
                           "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
              if (ok) break;
              __nanosleep(CGF_SLEEP_AUX_NS);
              (void)watchdog_expired(t0c, p.err, 10, polls);
            }
          };
          if (use == 0) wait_desc(0u);
          wait_desc(use + 1);
          CGF_EVENT(0, 1);
          mbar_wait<CGF_SLEEP_AUX_NS>(x_empty + stage, (use & 1) ^ 1, p.err, 2);
          CGF_EVENT(0, 2);
        }
        // ---- both CTAs: follow the descriptor of this tile
        const uint4 q0 = lds128(desc_addr(stage, use, 0u));
        const int tt0 = (int)q0.x, b0 = (int)q0.y;
        if (tt0 == -2) {                                   // STOP: wake the roles behind me, twice (one per pair)
          if (elect_one()) mbar_arrive(raw_full + stage);
          __syncwarp();
          if (++nstop == 2) break;
          continue;
        }
        const int fam = (int)lds32(desc_addr(stage, use, 0u) + 28);
        const uint4 r0 = lds128(desc_addr(stage, use, 1u));
        const int tt1 = (int)r0.x, b1 = (int)r0.y;
        const int head = fam / CBS;
        const int c_head = head * (KB * 64);
        const int nhalf = tt1 >= 0 ? 2 : 1;
        if (elect_one()) mbar_expect_tx(raw_full + stage, nhalf * (Cfg::kXStageBytes / 2 / CL));
        for (int hf = 0; hf < nhalf; ++hf) {
          const int tt = hf == 0 ? tt0 : tt1, b = hf == 0 ? b0 : b1;
          if (elect_one()) {
#pragma unroll
            for (int kbl = 0; kbl < KBL; ++kbl) {
              const int kb = (CL > 1 ? (int)crank * KBL : 0) + kbl;   // my channel half only
              tma_load_3d(sX + stage * Cfg::kXStageBytes + kb * Cfg::kXKBlock + hf * (kTile * 128), &tmap_x,
                          raw_full + stage, c_head + kb * 64, tt * kTile, b);
            }
          }
        }
        __syncwarp();
        // A tile of another head: its weights, AFTER its X loads -- the MMA warp learns the family from the
        // descriptor once x_full of this tile completes and only then releases the old weights (w_empty)
        if (head != cur_head) load_weights(head);
      }
      (void)nclusters;
    }
    __syncwarp();
  } else if (warp == kEpiWarps + 1) {
    // ===================================================== MMA issuer
    {
      const uint32_t w_lo0 = umma_desc_lo(sW), i_lo0 = umma_desc_lo(sI), x_lo0 = umma_desc_lo(sX);
      uint32_t witer = 0;
      int tn = 0; (void)tn;
      int cur_fam = -1;
      auto family_change = [&](int fam) {
        if (cur_fam >= 0) {                              // the old family's weights may be overwritten
          if (elect_one()) umma_commit(w_empty);
          __syncwarp();
        }
        cur_fam = fam;
        mbar_wait<CGF_SLEEP_AUX_NS>(w_full, witer & 1, p.err, 3);
        tc_fence_after();
        ++witer;
      };
      // the 40 MMAs of one tile of pair pr (= X stage pr): D_x, D_a, D_t
      auto issue_tile = [&](uint32_t pr, uint32_t use, int cb) {
        mbar_wait<CGF_SLEEP_AUX_NS>(t_empty + pr, (use & 1) ^ 1, p.err, 5);
        CGF_EVENT(1, 3);
        tc_fence_after();
        const uint32_t dcol = tmem_base + pr * kPairCols;
        const uint32_t x_lo = x_lo0 + pr * (Cfg::kXStageBytes >> 4);
        if (elect_one()) {
#pragma unroll
          for (int gate = 0; gate < 2; ++gate) {
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_bf16_lo(dcol + gate * kMmaN, w_lo0 + (((gate * KB + kb) * kKBlockBytes + k * 32) >> 4),
                             x_lo + ((kb * Cfg::kXKBlock + k * 32) >> 4), IDESC, (kb | k) != 0);
              }
            }
          }
          const uint32_t xt_lo = x_lo + cb * ((2 * Cfg::kXKBlock) >> 4);
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_bf16_lo(dcol + 2 * kMmaN, i_lo0 + ((kb * kKBlockBytes + k * 32) >> 4),
                           xt_lo + ((kb * Cfg::kXKBlock + k * 32) >> 4), IDESC, (kb | k) != 0);
            }
          }
          umma_commit(t_full + pr);
          if constexpr (CL > 1) umma_commit_mc(x_empty + pr, (uint16_t)((1u << CL) - 1u));   // my stage AND the peer's copy target
          else umma_commit(x_empty + pr);
        }
        __syncwarp();
        CGF_EVENT(1, 4);
      };
      if constexpr (!CONV) {
        uint32_t mq = 0;
        for (int sgi = 0; sgi < nsegs; ++sgi) {
          const Seg sg = sched.get(sgi);
          if (sg.count == 0) continue;
          const int fam = seg_family(sg);
          if (fam != cur_fam) family_change(fam);
          const int cb = fam % CBS;
#pragma unroll 1
          for (int m = 0; m < sg.count; ++m, ++mq) {
            const uint32_t pr = mq & 1u, use = mq >> 1;    // warpgroup pair == X stage
            CGF_EVENT(1, 1);
            mbar_wait<CGF_SLEEP_AUX_NS>(mma_ready + pr, use & 1, p.err, 4);
            CGF_EVENT(1, 2);
            issue_tile(pr, use, cb);
          }
        }
      } else {
        // dynamic schedule (see the producer): follow the descriptors until both pairs have seen STOP
        family_change((((int)blockIdx.x / CL) % (nfam / CBS)) * CBS + (int)crank);   // the producer's first load: the home family
        uint32_t nstop = 0;
#pragma unroll 1
        for (uint32_t mq = 0;; ++mq) {
          const uint32_t pr = mq & 1u, use = mq >> 1;
          CGF_EVENT(1, 1);
          mbar_wait<CGF_SLEEP_AUX_NS>(mma_ready + pr, use & 1, p.err, 4);   // x_full: the convolution warps have seen the tile
          CGF_EVENT(1, 2);
          const uint32_t da = desc_addr(pr, use, 0u);
          if ((int)lds32(da) == -2) {
            if (++nstop == 2) break;
            continue;
          }
          const int fam = (int)lds32(da + 28);
          if (fam != cur_fam) family_change(fam);
          issue_tile(pr, use, fam % CBS);
        }
      }
    }
    __syncwarp();
  } else if (warp >= kEpiWarps + 2) {
    // warps 18 / 19 complete the fifth warpgroup (registers are allocated per four warps) and idle.
    // (Round 2's first version ran the in-kernel convolution here: two warps on two of the SM's four
    // sub-partitions, the whole head's 256 channels per CTA -- 165-175 us at config 2 against 136 us for
    // the two-kernel route.  It now runs in the epilogue warpgroups, see conv_tile below.)
  } else {
    // ===================================================== epilogue warpgroups
    // One warpgroup, tile after tile:
    //   request the state word of tile k-1's predecessor, wait for tile k's
    //   accumulators (the look-back round trip hides behind that wait);
    //   F(k-1) finish the previous tile: carry-in by decoupled look-back and the
    //          replay pass that writes y;
    //   G(k)   gates: accumulators -> registers (rounded to bf16) -> (a, x~) into
    //          my TMEM state columns; the slot goes back to the MMA warp half way
    //          through, so the MMAs of the pair's next tile run under G and F;
    //          publish the tile's aggregate.
    // The launch is a programmatic dependent of the prologue kernel (epoch bump,
    // -8*softplus, reset bitmask): barrier / TMEM set-up above and the producer's
    // weight and X loads overlap the prologue; its outputs are read only below.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int wg = warp >> 2;
    const uint32_t pr = wg >> 1, hf = wg & 1;
    const int chl = (warp & 3) * 32 + lane;           // TMEM lane = channel inside the column
    // TMEM addresses of this thread: accumulators of my half, my state columns
    const uint32_t tm_acc = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + pr * kPairCols + hf * kTile;
    const uint32_t tm_state = tm_acc - hf * kTile + 3 * kMmaN + hf * kTile;
    const unsigned epoch = *p.epoch;
    uint32_t mq = 0;
    int tn = 0; (void)tn;
    const int trole = 2 + wg; (void)trole;
    const bool twarp = (warp & 3) == 0; (void)twarp;

    // the tile whose pass 2 is still owed
    struct Pending { bool on; int tt, b, nvalid, ch; size_t widx; uint16_t* yp; float P, H; };
    Pending pd{}; pd.on = false;
    const size_t wstep = (size_t)p.B * kMch;             // exchange words: one time tile back

    auto release_slot = [&]() {                          // my half of the accumulators is consumed
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty + pr);
    };

    // F: carry chain + pass 2 of the pending tile
    // carry-in of the pending tile by decoupled look-back; publishes its state
    auto resolve_carry = [&](unsigned long long early) -> float {
      // `early`: the predecessor's state word, requested before the accumulator wait
      const int tt = pd.tt;
      float c0;
      if (tt == 0 || (CGF_ABLATE & 1)) {
        c0 = p.h0 != nullptr ? p.h0[(size_t)pd.b * p.E + pd.ch] : 0.0f;
      } else if (__all_sync(0xffffffffu, (unsigned)early == epoch)) {
        c0 = tagged_value(early);
      } else {
        // Walk back over tiles whose aggregate is published until one with a
        // published state is found (kLook tiles per round trip), then apply the
        // aggregates passed on the way one by one, left to right -- exactly the
        // expression each of those tiles evaluates for its own state, so the
        // carry does not depend on timing (bit-reproducible).
        constexpr int kLook = LOOK;
        int end = tt;                                    // tiles [end, tt): aggregate seen
        const long long t_start = clock64();
        unsigned polls = 0;
        c0 = 0.0f;
        // aggregates of the two nearest predecessors, as accepted by the walk: the fold below uses them
        // instead of loading them again (one round trip less on every hop of the carry chain)
        // (in the depth-2 instances only, i.e. small batches, where the chain is the bound: B=2, T=8192
        // 135.9 -> 132.2 us; at B >= 8 it is 1 % slower: profiles/r3_ab_keep_aggregates.txt)
        constexpr bool kKeepAgg = CGF_KEEP_AGG && LOOK == 2;
        float keepP[2] = {1.0f, 1.0f}, keepH[2] = {0.0f, 0.0f};
        for (;;) {
          const int depth = end < kLook ? end : kLook;
          unsigned long long wpf[kLook], wap[kLook], wah[kLook];
#pragma unroll
          for (int k = 0; k < kLook; ++k) {
            if (k < depth) {
              const size_t src = pd.widx - (size_t)(tt - end + k + 1) * wstep;
              wpf[k] = ld_relaxed_u64(p.pref + src);
              if (k > 0 || kLook > 1) {
                wap[k] = ld_relaxed_u64(p.agg_p + src);
                wah[k] = ld_relaxed_u64(p.agg_h + src);
              }
            } else {
              wpf[k] = wap[k] = wah[k] = 0ull;
            }
          }
          bool done = false;
          int used = 0;                                  // aggregates accepted this round
#pragma unroll
          for (int k = 0; k < kLook; ++k) {
            if (!done && k == used && k < depth) {
              if (__all_sync(0xffffffffu, (unsigned)wpf[k] == epoch)) {
                c0 = tagged_value(wpf[k]);
                done = true;
              } else {
                if (kLook == 1) {                        // second round trip only when needed
                  const size_t src = pd.widx - (size_t)(tt - end + 1) * wstep;
                  wap[0] = ld_relaxed_u64(p.agg_p + src);
                  wah[0] = ld_relaxed_u64(p.agg_h + src);
                }
                if (__all_sync(0xffffffffu, (unsigned)wap[k] == epoch && (unsigned)wah[k] == epoch)) {
                  used = k + 1;
                  const int dist = tt - end + k + 1;     // tile tt - dist
                  if (kKeepAgg && dist <= 2) { keepP[dist - 1] = tagged_value(wap[k]); keepH[dist - 1] = tagged_value(wah[k]); }
                }
              }
            }
          }
          end -= used;                                   // (tile 0 never publishes an aggregate: end stays > 0)
          if (done) break;
          if (used == 0) __nanosleep(40);
          polls += 3;
          if (__any_sync(0xffffffffu, watchdog_expired(t_start, p.err, 7, polls))) break;
        }
        for (int k = end; k < tt; ++k) {                 // state(end-1) -> ... -> state(tt-1)
          const int dist = tt - k;
          if (kKeepAgg && dist <= 2) {
            c0 = fmaf(dist == 2 ? keepP[1] : keepP[0], c0, dist == 2 ? keepH[1] : keepH[0]);
          } else {
            const size_t src = pd.widx - (size_t)dist * wstep;
            c0 = fmaf(tagged_value(ld_relaxed_u64(p.agg_p + src)), c0,
                      tagged_value(ld_relaxed_u64(p.agg_h + src)));
          }
        }
      }
      if (tt + 1 < p.ntt) st_relaxed_u64(p.pref + pd.widx, pack_tagged(fmaf(pd.P, c0, pd.H), epoch));
      if (twarp) CGF_EVENT(trole, 5);
      return c0;
    };
    // replay of 8 steps of the pending tile: h = a*h + x~ from the true carry, mul
    // then add as the reference loop (:196); y leaves as bf16 (2-byte stores: a
    // warp writes 64 contiguous bytes per row)
    auto replay_chunk = [&](int c, const uint32_t (&st)[8], float& h, const uint32_t* gm) {
      const int E = rowE;
      uint32_t o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t a2 = st[i], n2 = st[4 + i];
        const float y0 = add_bf_lo(n2, bf_lo(a2) * h);
        const float y1 = add_bf_hi(n2, bf_hi(a2) * y0);
        h = y1;
        o[i] = pack_bf2(y0, y1);
        if constexpr (MUL) o[i] = bf2_mul(o[i], gm[2 * i] | (gm[2 * i + 1] << 16));   // r(r(h) * gate), :651
      }
      uint16_t* yc = pd.yp + (size_t)(c * 8) * E;
      if ((CGF_ABLATE & 2) && o[0] != 0x12345678u) return;
      // (lane pairs swapping halves through a shuffle so that every lane writes ONE 4-byte word per pair
      // of steps -- half the stores and address arithmetic -- was measured SLOWER: 126.7 vs 125.4 us at
      // config 2, 879 vs 873 us at B=16, T=8192; profiles/r3_ab_look_shfl.txt)
      if (c * 8 + 8 <= pd.nvalid) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          st_u16(yc + (2 * i) * E, o[i]);
          st_u16(yc + (2 * i + 1) * E, o[i] >> 16);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (c * 8 + 2 * i < pd.nvalid) st_u16(yc + (2 * i) * E, o[i]);
          if (c * 8 + 2 * i + 1 < pd.nvalid) st_u16(yc + (2 * i + 1) * E, o[i] >> 16);
        }
      }
    };
    auto retire_pending = [&](float h) {                 // the pending tile is complete
      if (p.last_h != nullptr && pd.tt == p.ntt - 1) p.last_h[(size_t)pd.b * p.E + pd.ch] = h;
      pd.on = false;
      if (twarp) CGF_EVENT(trole, 6);
    };
    // F: carry chain + the whole replay pass of the pending tile
    auto finish = [&](unsigned long long early) {
      const int E = rowE;
      // gating-product operand of the first 8 steps, requested before the
      // look-back so that its latency hides behind it
      uint32_t gm[8], gn[8];
      const uint16_t* gmp = nullptr;
      if constexpr (MUL) {
        gmp = p.gate_mul + (pd.yp - p.y);
#pragma unroll
        for (int i = 0; i < 8; ++i) gm[i] = i < pd.nvalid ? ld_u16(gmp + i * E) : 0u;
      }
      float h = resolve_carry(early);
#pragma unroll 1
      for (int c = 0; c < ((CGF_ABLATE & 4) ? 0 : kTile / 8); ++c) {
        uint32_t st[8];
        tmem_ld8(tm_state + c * 8, st);
        if constexpr (MUL) {                             // operand of chunk c + 1, one chunk ahead
#pragma unroll
          for (int i = 0; i < 8; ++i)
            gn[i] = (c + 1 < kTile / 8 && (c + 1) * 8 + i < pd.nvalid) ? ld_u16(gmp + ((c + 1) * 8 + i) * E) : 0u;
        }
        tmem_wait_ld();
        replay_chunk(c, st, h, gm);
        if constexpr (MUL) {
#pragma unroll
          for (int i = 0; i < 8; ++i) gm[i] = gn[i];
        }
      }
      retire_pending(h);
    };
    auto request_pred = [&]() -> unsigned long long {    // state word of the pending tile's predecessor
      return (pd.on && pd.tt > 0) ? ld_relaxed_u64(p.pref + pd.widx - wstep) : 0ull;
    };

    auto read_desc = [&](uint32_t use) -> TileDesc {         // my half of pair pr's MMA tile number `use`
      const uint32_t da = desc_addr(pr, use, hf);
      const uint4 q0 = lds128(da), q1 = lds128(da + 16);
      return TileDesc{(int)q0.x, (int)q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, (int)q1.w};
    };
    int cur_fam = -1, ch = 0;
    uint32_t bx2 = 0, ba2 = 0, sp2 = 0;
    auto load_family = [&](int fam) {                      // per-channel constants of a new family
      if (fam == cur_fam) return;
      cur_fam = fam;
      ch = fam * kMch + chl;
      bx2 = 0; ba2 = 0;
      if (p.bias_x != nullptr) { const uint32_t v = p.bias_x[ch]; bx2 = v | (v << 16); }
      if (p.bias_a != nullptr) { const uint32_t v = p.bias_a[ch]; ba2 = v | (v << 16); }
      sp2 = p.neg8sp_bf[ch]; sp2 |= sp2 << 16;
    };
    // G(k): gates of one scan tile out of my half of the accumulators, aggregate published, tile queued for F
    auto tile_body = [&](const TileDesc& td) {
      // (descriptor fields are the same in every lane; the shuffles tell the compiler so)
      const int tt = td.tt, b = td.b;
      const unsigned rbits = td.rbits;
      const int t0 = tt * kTile;
      const int nvalid = p.T - t0;                         // >= 1; >= kTile for a full tile
      const bool fast_tile = CGF_PRELOAD && rbits == 0u && nvalid >= kTile;
      float P = 1.0f, Hh = 0.0f;
      // gates for one bf16x2 pair of steps (t, t+1) -> (a, x~); the tile's
      // transform h -> P*h + H is accumulated on the way
      auto gate_step = [&](uint32_t xc, uint32_t gxr, uint32_t gar, int tl, auto slow_tag,
                           uint32_t& a2, uint32_t& n2) {
        constexpr bool SLOW = decltype(slow_tag)::value;
        gate_pair_emul<FAST, false>(xc, gxr, gar, bx2, ba2, sp2, a2, n2);
        if constexpr (SLOW) {                            // tl = step of the low half inside the tile
          const unsigned r2 = (rbits >> tl) & 3u;
          if (r2 != 0u) {                                // document start inside the pair
            uint32_t az, nr;
            gate_pair_emul<FAST, true>(xc, gxr, gar, bx2, ba2, sp2, az, nr);
            if (r2 & 1u) { a2 &= 0xffff0000u; n2 = (n2 & 0xffff0000u) | (nr & 0x0000ffffu); }
            if (r2 & 2u) { a2 &= 0x0000ffffu; n2 = (n2 & 0x0000ffffu) | (nr & 0xffff0000u); }
          }
          if (tl + 1 >= nvalid) {                        // steps beyond T are identities
            if (tl >= nvalid) { a2 = kOne2; n2 = 0u; }
            else { a2 = (a2 & 0x0000ffffu) | 0x3f800000u; n2 &= 0x0000ffffu; }
          }
        }
        if constexpr (DBG && !CGF_TRACE) {
          const int tg = t0 + tl;
          const size_t plane = (size_t)p.B * p.T * p.E;
          const size_t o0 = ((size_t)b * p.T + tg) * p.E + ch;
          if (tg < p.T) {
            p.dbg[o0] = (uint16_t)(gxr & 0xffffu); p.dbg[plane + o0] = (uint16_t)(gar & 0xffffu);
            p.dbg[2 * plane + o0] = (uint16_t)(xc & 0xffffu);
          }
          if (tg + 1 < p.T) {
            p.dbg[o0 + p.E] = (uint16_t)(gxr >> 16); p.dbg[plane + o0 + p.E] = (uint16_t)(gar >> 16);
            p.dbg[2 * plane + o0 + p.E] = (uint16_t)(xc >> 16);
          }
        }
        const float al = bf_lo(a2), ah = bf_hi(a2);
        Hh = add_bf_lo(n2, al * Hh);                     // mul then add, as the reference loop (:196)
        Hh = add_bf_hi(n2, ah * Hh);
        P *= al; P *= ah;
      };
      if (fast_tile) {
        // ---- common case, half a tile (8 pairs) at a time: pull the
        // accumulators out of TMEM, rounded to bf16 on the way (the GEMM output
        // the reference materialises, :136-142), then the gate math from
        // registers.  The slot goes back to the MMA warp as soon as the second
        // half has been read.
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t gx[8], ga[8], xv[8];
          {
            uint32_t dx[16], da[16], dt[16];
            const uint32_t col = tm_acc + hh * 16;
            tmem_ld8(col, *reinterpret_cast<uint32_t(*)[8]>(&dx[0]));
            tmem_ld8(col + 8, *reinterpret_cast<uint32_t(*)[8]>(&dx[8]));
            tmem_ld8(col + kMmaN, *reinterpret_cast<uint32_t(*)[8]>(&da[0]));
            tmem_ld8(col + kMmaN + 8, *reinterpret_cast<uint32_t(*)[8]>(&da[8]));
            tmem_ld8(col + 2 * kMmaN, *reinterpret_cast<uint32_t(*)[8]>(&dt[0]));
            tmem_ld8(col + 2 * kMmaN + 8, *reinterpret_cast<uint32_t(*)[8]>(&dt[8]));
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              gx[i] = pack_bf2(__uint_as_float(dx[2 * i]), __uint_as_float(dx[2 * i + 1]));
              ga[i] = pack_bf2(__uint_as_float(da[2 * i]), __uint_as_float(da[2 * i + 1]));
              xv[i] = pack_bf2(__uint_as_float(dt[2 * i]), __uint_as_float(dt[2 * i + 1]));
            }
          }
          if (hh == 1) {
            release_slot();
            if (twarp) CGF_EVENT(trole, 3);
          }
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t st[8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
              gate_step(xv[c * 4 + i], gx[c * 4 + i], ga[c * 4 + i], hh * 16 + c * 8 + 2 * i, FalseTag{}, st[i], st[4 + i]);
            tmem_st8(tm_state + hh * 16 + c * 8, st);
          }
        }
      } else {
        // document starts or a ragged tail inside the tile (rare, warp-uniform):
        // gates chunk by chunk out of TMEM
#pragma unroll 1
        for (int c = 0; c < kTile / 8; ++c) {
          uint32_t dx[8], da[8], dt[8], st[8];
          tmem_ld8(tm_acc + c * 8, dx);
          tmem_ld8(tm_acc + kMmaN + c * 8, da);
          tmem_ld8(tm_acc + 2 * kMmaN + c * 8, dt);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t gxr = pack_bf2(__uint_as_float(dx[2 * i]), __uint_as_float(dx[2 * i + 1]));
            const uint32_t gar = pack_bf2(__uint_as_float(da[2 * i]), __uint_as_float(da[2 * i + 1]));
            const uint32_t xc = pack_bf2(__uint_as_float(dt[2 * i]), __uint_as_float(dt[2 * i + 1]));
            gate_step(xc, gxr, gar, c * 8 + 2 * i, TrueTag{}, st[i], st[4 + i]);
          }
          tmem_st8(tm_state + c * 8, st);
        }
        release_slot();
      }
      tmem_wait_st();
      if (twarp) CGF_EVENT(trole, 4);
      // publish the tile's aggregate (tile 0 publishes its state right away in
      // F instead) and queue the tile for F
      pd.on = true; pd.tt = tt; pd.b = b; pd.nvalid = nvalid; pd.ch = ch; pd.P = P; pd.H = Hh;
      pd.widx = (size_t)td.widx + chl;
      pd.yp = p.y + (((size_t)td.eoff_hi << 32) | td.eoff_lo) + ch;
      if (tt > 0 && tt + 1 < p.ntt) {
        st_relaxed_u64(p.agg_p + pd.widx, pack_tagged(P, epoch));
        st_relaxed_u64(p.agg_h + pd.widx, pack_tagged(Hh, epoch));
      }
      if (twarp) CGF_EVENT(trole, 7);
    };

    if constexpr (!CONV) {
      for (int sgi = 0; sgi < nsegs; ++sgi) {
        const Seg sg = sched.get(sgi);
        if (sg.count == 0) continue;
        const int fam = sg.fam;
        load_family(fam);
#pragma unroll 1
        for (int m = 0; m < sg.count; ++m, ++mq) {
          if ((mq & 1u) != pr) continue;
          const uint32_t use = mq >> 1;
          if (twarp) CGF_EVENT(trole, 8);
          const unsigned long long early = request_pred();
          if (twarp) CGF_EVENT(trole, 1);
          mbar_wait(t_full + pr, use & 1, p.err, 6);
          if (twarp) CGF_EVENT(trole, 2);
          tc_fence_after();
          const TileDesc td = read_desc(use);              // written by the producer before the TMA loads of this tile
          if (td.tt < 0) {                                 // odd tile count: nothing in my half
            release_slot();
            if (pd.on) finish(early);
            continue;
          }
          // F(k-1): before the new tile
          if (pd.on) finish(early);
          tile_body(td);
        }
      }
    } else {
      // ------------------------------------------------------------------------------------------
      // CONV: this warpgroup also convolves ITS half (32 rows) of its pair's NEXT X stage, in place,
      // between F(k-1) and G(k) -- the convolution's instructions are spread evenly over the SM's four
      // sub-partitions (a warpgroup has one warp on each) and the TMA round trip of the next tile's raw
      // rows hides behind F.  128 threads = 64 bf16x2 columns (this CTA's 128 input channels: warp ->
      // K block, lane -> 4-byte word of the 128-byte row, conflict-free) x 2 row segments of 16 rows.
      // A thread reads its 16 rows and its three halo rows (segment 0: the rows before the tile, straight
      // from global memory, requested before the wait for the TMA data; segment 1: rows 13-15 of the
      // half) BEFORE a warpgroup barrier and writes after it, so in-place is safe.  Arithmetic = the
      // reference's accumulation order with one bf16 rounding per eager op (layers.py:530-536), bit-exact
      // with cg::conv1d_w4_kernel.  CL == 2: the convolved half is then sent to the peer CTA (one bulk
      // shared -> shared::cluster copy per K block, completing on the PEER's x_full), which needs it as
      // the other half of the K range of its gate GEMMs -- the convolution is computed once per head.
      // ------------------------------------------------------------------------------------------
      const int wq = warp & 3;
      const int cseg = wq >> 1;                            // rows 16 * cseg .. + 15 of my half
      const int ckb = (CL > 1 ? (int)crank * KBL : 0) + (wq & 1);   // my K block of the stage
      const uint32_t lofs = (uint32_t)lane << 2;           // my word of a 128-byte row (16-byte chunk lane >> 2)
      const uint32_t peer = crank ^ 1u;
      // The taps of the 128 input channels of this CTA's HOME head sit in shared memory, written once by
      // warpgroup 0; tiles of other heads (drawn once the home head is exhausted) take theirs from global
      // memory per tile.
      const int home_head = ((int)blockIdx.x / CL) % (nfam / CBS);
      {
        if (wg == 0) {
          const int cht = home_head * (KB * 64) + (CL > 1 ? (int)crank * 128 : 0) + (int)threadIdx.x;   // threads 0..127
#pragma unroll
          for (int k = 0; k < 4; ++k)
            asm volatile("st.shared.u16 [%0], %1;" :: "r"(sTap + k * 256 + threadIdx.x * 2), "h"(p.conv_w[(size_t)k * p.E + cht]));
          asm volatile("st.shared.u16 [%0], %1;" :: "r"(sTap + 1024 + threadIdx.x * 2), "h"(p.conv_b[cht]));
        }
        named_bar_sync(1, kEpiWarps * 32);
      }
      const uint32_t row0 = sX + pr * Cfg::kXStageBytes + (uint32_t)ckb * Cfg::kXKBlock +
                            (uint32_t)(hf * kTile + cseg * 16) * 128u;   // my 16 rows of the stage; (row & 7) == (j & 7)
      const uint32_t tap_a = sTap + (uint32_t)((wq & 1) * 32 + lane) * 4u;
      // The three rows before my segment of a tile come straight from global memory (L2: the TMA reads
      // them at about the same time); x[t < 0] = 0 (layers.py:484-492).  They are requested a whole tile
      // ahead -- the descriptor of the pair's next tile is readable as soon as this warpgroup has waited on
      // raw_full for the current one (TileDesc) -- so the round trip never shows.
      uint32_t hq1 = 0u, hq2 = 0u, hq3 = 0u;               // x[ts-1], x[ts-2], x[ts-3] of the next convolution
      auto halo_request = [&](uint32_t use_n) {
        const uint32_t da = desc_addr(pr, use_n, hf);
        const int tt_n = (int)lds32(da);
        hq1 = 0u; hq2 = 0u; hq3 = 0u;
        if (tt_n >= 0) {
          const uint32_t e_lo = lds32(da + 16), e_hi = lds32(da + 20);
          const int head_n = (int)lds32(da + 28) / CBS;
          const int ts = tt_n * kTile + cseg * 16;         // first step of my segment
          const int chp = head_n * (KB * 64) + ckb * 64 + lane * 2;
          const uint16_t* xb = p.x_lin + (((size_t)e_hi << 32) | e_lo) + (size_t)(cseg * 16) * rowE + chp;
          if (ts >= 3 && ts <= p.T) {                      // interior (the common case)
            const uint16_t* x3 = xb - 3 * (size_t)rowE;
            hq3 = ldg32_nc(x3); hq2 = ldg32_nc(x3 + rowE); hq1 = ldg32_nc(x3 + 2 * (size_t)rowE);
          } else {
            if (ts >= 1 && ts - 1 < p.T) hq1 = ldg32_nc(xb - (size_t)p.E);
            if (ts >= 2 && ts - 2 < p.T) hq2 = ldg32_nc(xb - 2 * (size_t)p.E);
            if (ts >= 3 && ts - 3 < p.T) hq3 = ldg32_nc(xb - 3 * (size_t)p.E);
          }
        }
      };
      // convolves my half of pair pr's tile number `use`; false: that tile is the STOP marker
      auto conv_tile = [&](const uint32_t use) -> bool {
        if (twarp) CGF_EVENT(trole, 9);
        mbar_wait(raw_full + pr, use & 1, p.err, 9);       // the raw rows of the tile
        if (twarp) CGF_EVENT(trole, 10);
        const TileDesc td = read_desc(use);
        const bool valid = td.tt >= 0;                     // uniform over the warpgroup
        if (valid && !(CGF_ABLATE & 8)) {
          const int head = td.fam / CBS;
          const int chp = head * (KB * 64) + ckb * 64 + lane * 2;   // my two channels
          const uint32_t h1 = hq1, h2 = hq2, h3 = hq3;
          // a thread reads and writes only its own 16 words of the stage: no hazard, no barrier
          uint32_t xr[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) xr[j] = lds32(row0 + j * 128 + (lofs ^ ((uint32_t)(j & 7) << 4)));
          uint32_t k0, k1, k2, k3, kb_;                    // w[k] multiplies x[t - (3 - k)] (layers.py:530)
          if (head == home_head) {
            k0 = lds32(tap_a); k1 = lds32(tap_a + 256); k2 = lds32(tap_a + 512); k3 = lds32(tap_a + 768); kb_ = lds32(tap_a + 1024);
          } else {                                         // several families per CTA (small grids): per tile
            k0 = ldg32_nc(p.conv_w + chp);
            k1 = ldg32_nc(p.conv_w + (size_t)p.E + chp);
            k2 = ldg32_nc(p.conv_w + 2 * (size_t)p.E + chp);
            k3 = ldg32_nc(p.conv_w + 3 * (size_t)p.E + chp);
            kb_ = ldg32_nc(p.conv_b + chp);
          }
          // bit r of nzw: segment_pos[t0 + r - 2] != 0
          const unsigned long long nzw = ((unsigned long long)(~td.rbits) << 2) | (unsigned long long)((~td.rprev) >> 30);
          const bool upstream = p.mask_mode != 0;
          // no document start near the tile: every tap is live (the common case)
          const bool plain = upstream ? (nzw & 0x3ffffffffull) == 0x3ffffffffull
                                      : (nzw & 0xffffffffull) == 0xffffffffull;
          if (plain) {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
              const int j = jj < 13 ? jj + 3 : jj - 13;    // rows 3..15, then the rows that need the halo
              const uint32_t a1 = j >= 1 ? xr[j - 1] : h1;
              const uint32_t a2 = j >= 2 ? xr[j - 2] : (j == 1 ? h1 : h2);
              const uint32_t a3 = j >= 3 ? xr[j - 3] : (j == 2 ? h1 : (j == 1 ? h2 : h3));
              uint32_t acc = bf2_mul(xr[j], k3);                   // shift 0
              acc = bf2_add(acc, bf2_mul(a1, k2));                 // shift 1
              acc = bf2_add(acc, bf2_mul(a2, k1));                 // shift 2
              acc = bf2_add(acc, bf2_mul(a3, k0));                 // shift 3
              sts32(row0 + j * 128 + (lofs ^ ((uint32_t)(j & 7) << 4)), bf2_add(acc, kb_));   // + b, :536
            }
          } else {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
              const int j = jj < 13 ? jj + 3 : jj - 13;
              const uint32_t a1 = j >= 1 ? xr[j - 1] : h1;
              const uint32_t a2 = j >= 2 ? xr[j - 2] : (j == 1 ? h1 : h2);
              const uint32_t a3 = j >= 3 ? xr[j - 3] : (j == 2 ? h1 : (j == 1 ? h2 : h3));
              const unsigned w3 = (unsigned)(nzw >> (cseg * 16 + j)) & 7u;   // bit 0: seg[t-2], 1: seg[t-1], 2: seg[t]  (!= 0)
              bool m1 = true, m2 = true, m3;
              if (!upstream) {
                m3 = (w3 & 1u) != 0;                       // fork: only seg[t-2], layers.py:629-632
              } else {
                m1 = (w3 & 4u) != 0;                       // upstream: seg[t-s+1 .. t] all != 0
                m2 = (w3 & 6u) == 6u;
                m3 = (w3 & 7u) == 7u;
              }
              uint32_t acc = bf2_mul(xr[j], k3);
              acc = bf2_add(acc, bf2_mul(m1 ? a1 : 0u, k2));
              acc = bf2_add(acc, bf2_mul(m2 ? a2 : 0u, k1));
              acc = bf2_add(acc, bf2_mul(m3 ? a3 : 0u, k0));
              sts32(row0 + j * 128 + (lofs ^ ((uint32_t)(j & 7) << 4)), bf2_add(acc, kb_));
            }
          }
          // the convolution's returned cache: the last three INPUT rows of the sequence, left zero
          // padded (layers.py:542-543, :650-662)
          if (p.conv_cache != nullptr && td.tt == p.ntt - 1 && cseg == 0) {
            const uint16_t* xq = p.x_lin + ((size_t)td.b * p.T) * p.E + chp;
            uint16_t* cb = p.conv_cache + ((size_t)td.b * 3) * p.E + chp;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              const int ti = p.T - 3 + r;
              *reinterpret_cast<uint32_t*>(cb + (size_t)r * p.E) = ti >= 0 ? ldg32_nc(xq + (size_t)ti * p.E) : 0u;
            }
          }
        }
        // my warp's stores (generic proxy) before the MMA's operand reads and the bulk copy's source reads
        // (async proxy).  Every warp hands over ITS 16 rows x 128 B on its own: no warpgroup barrier.
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (CL > 1 && valid && !(CGF_ABLATE & 16)) {
            bulk_copy_to_peer(mapa_u32(row0, peer), row0, 16u * 128u, mapa_u32(smem_u32(x_full + pr), peer));
            // my arrival, and the bytes the peer's same warp sends into MY stage
            mbar_expect_tx(x_full + pr, 16u * 128u);
          } else {
            mbar_arrive(x_full + pr);
          }
        }
        if (twarp) CGF_EVENT(trole, 11);
        return td.tt != -2;
      };

      // prologue: the first tile of my pair (its descriptor is written right before the arrive on raw_full)
      mbar_wait(raw_full + pr, 0u, p.err, 9);
      halo_request(0u);
      bool alive = conv_tile(0u);
#pragma unroll 1
      for (uint32_t use = 0; alive; ++use) {
        if (twarp) CGF_EVENT(trole, 8);
        const unsigned long long early = request_pred();
        // halo rows of the NEXT convolution: its descriptor became readable with raw_full of this tile
        halo_request(use + 1);
        if (twarp) CGF_EVENT(trole, 1);
        mbar_wait(t_full + pr, use & 1, p.err, 6);
        if (twarp) CGF_EVENT(trole, 2);
        tc_fence_after();
        const bool mine = (int)lds32(desc_addr(pr, use, hf)) >= 0;   // odd tile count: nothing in my half of the last pair
        if (!mine) release_slot();
        if (pd.on) finish(early);                          // F(k-1)
        alive = conv_tile(use + 1);                        // the raw rows of my pair's next tile have landed under F
        if (mine) {                                        // G(k)  (the slot is rewritten four uses later)
          const TileDesc td = read_desc(use);
          load_family(td.fam);
          tile_body(td);
        }
      }
    }
    if (pd.on) finish(request_pred());                     // the last tile of this warpgroup
  }

#if CGF_TRACE
  // per-CTA end time (ns, %globaltimer) behind the role time lines: [7 * 1024 + blockIdx.x]
  if (threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    reinterpret_cast<unsigned long long*>(p.dbg)[7 * 1024 + blockIdx.x] = t;
  }
#endif
  // teardown: every role is done with TMEM; a CTA of a cluster must not exit while its peer may still
  // signal its barriers (the multicast commit of the peer's last MMAs)
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();
  if (warp == kEpiWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace fused
}  // namespace cg
