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
// Roles (576 threads, 1 CTA / SM, persistent):
//   warps 0-15   four epilogue warpgroups in two pairs; pair p owns TMEM columns
//                [256p, 256p+256): 192 accumulator + 2 x 32 state.  Warpgroup
//                (p, h) works on half h (columns 32h..32h+31) of every MMA tile
//                issued for pair p; MMA tiles alternate between the pairs.
//   warp 16      TMA producer: packed, pre-swizzled gate weights once per
//                family, then two X boxes [32 steps x head width] per MMA tile
//                through a 3-D tensor map (SWIZZLE_128B, zero fill beyond T)
//   warp 17      MMA issuer (one thread), tcgen05.commit -> mbarriers
// CTA i works on family i % families; the CTAs of one family take its tiles
// two at a time, round-robin in time-major order, so look-back dependencies
// always point to tiles that are already running (all CTAs are co-resident:
// grid <= #SMs, 1 CTA / SM).
#pragma once

#include <cuda.h>

#include "cg_common.cuh"
#include "cg_scan.cuh"

// tuning switches (scripts/build_variants.py); the defaults are the shipped configuration
#ifndef CGF_SLEEP_NS
#define CGF_SLEEP_NS 64
#endif
#ifndef CGF_PRELOAD
#define CGF_PRELOAD 1
#endif
// look-back depth per round trip: 2 = the predecessor's state word plus the aggregates /
// state of the tile before it in ONE round trip.  Measured (us, fused kernel alone):
// B=2,T=8192 141.6 -> 129.4; B=8,T=2048 107.0 -> 107.0; B=32,T=768 144.3 -> 144.0-145.9;
// depth 4: 131.4 / 111.1 / 146.0 (more loads per poll than the chain is long)
#ifndef CGF_LOOK
#define CGF_LOOK 2
#endif
#ifndef CGF_HINT_NS
#define CGF_HINT_NS 20000
#endif
// CGF_INTERLEAVE: the replay pass of the previous tile runs inside the gate loop
// of the current one (same basic block: the compiler fills the gate math's MUFU
// latencies with the replay's FMA / store work) instead of before it
#ifndef CGF_LATE_WAIT_ST
#define CGF_LATE_WAIT_ST 0
#endif
#ifndef CGF_WAIT_LATE
#define CGF_WAIT_LATE 0
#endif
#ifndef CGF_INTERLEAVE
#define CGF_INTERLEAVE 0   // measured: no gain (110.8 vs 109.8 us), the kernel is throughput-bound
#endif
static_assert(!(CGF_LATE_WAIT_ST && CGF_INTERLEAVE), "the woven replay reads the state columns without that wait");
// CGF_ABLATE (timing experiments only, results are WRONG): 1 = no look-back,
// 2 = no y stores, 4 = no replay pass at all
#ifndef CGF_ABLATE
#define CGF_ABLATE 0
#endif
// CGF_TRACE: debug_out becomes a timeline buffer [6 roles][1024] of
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
  int words;                       // words per row (= ntt: one word per 32-step tile)
  const float* h0;                 // [B,E] or null
  uint16_t* y;                     // [B,T,E]
  float* last_h;                   // [B,E] or null
  const unsigned* epoch;
  unsigned long long* agg_p;       // [families][ntt][B][128]
  unsigned long long* agg_h;
  unsigned long long* pref;
  const int* conv_flags;           // optional [ceil(T/64)][B]: channel tiles of x the Conv1D producer kernel has finished
  int conv_need;                   // ... out of this many (E / 64)
  const uint16_t* gate_mul;        // optional [B,T,E]: y <- round_bf16(y * gate_mul) (RecurrentBlock, modules.py:651)
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
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"((unsigned)CGF_HINT_NS) : "memory");   // suspend-time hint (ns)
  return ok != 0;
}
// Bounded wait: a protocol bug must end the launch, never hang the device.  On
// a timeout the wait raises the error flag (workspace header) and returns; once
// the flag is up every later wait returns at once, so the kernel drains (with
// garbage results) and the host finds the flag.
constexpr long long kWatchdogCycles = 1000000000LL;   // ~0.5 s
__device__ __forceinline__ bool watchdog_expired(long long t0, int* err, int code, unsigned& polls) {
  if ((++polls & 63u) != 0u) return false;
  if (*reinterpret_cast<volatile int*>(err) != 0) return true;
  if (clock64() - t0 > kWatchdogCycles) { atomicCAS(err, 0, code); return true; }
  return false;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err, int code) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  unsigned polls = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (CGF_SLEEP_NS > 0) __nanosleep(CGF_SLEEP_NS);   // waiting warps must not eat the issue slots of the working ones
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
// y stores.  CGF_YST: 0 = default policy, 1 = .cs (streaming / evict-first), 2 = .wt
// (write-through), 3 = L2::evict_first cache-hint policy.  The 84 MB of y a
// config-2 launch writes would otherwise sit dirty in the 126 MB L2 and be
// written back under the NEXT kernel, which then pays for it.
#ifndef CGF_YST
#define CGF_YST 0   // measured: no policy changes the step time (149-153 us for all four)
#endif
__device__ __forceinline__ void st_u16(uint16_t* p, uint32_t v) {
#if CGF_YST == 1
  asm volatile("st.global.cs.u16 [%0], %1;" :: "l"(p), "h"(static_cast<uint16_t>(v)) : "memory");
#elif CGF_YST == 2
  asm volatile("st.global.wt.u16 [%0], %1;" :: "l"(p), "h"(static_cast<uint16_t>(v)) : "memory");
#elif CGF_YST == 3
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("st.global.L2::cache_hint.u16 [%0], %1, %2;" :: "l"(p), "h"(static_cast<uint16_t>(v)), "l"(pol) : "memory");
#else
  asm volatile("st.global.u16 [%0], %1;" :: "l"(p), "h"(static_cast<uint16_t>(v)) : "memory");
#endif
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

template <int KB>
struct FusedCfg {
  static constexpr uint32_t kWBytes = 2u * KB * kKBlockBytes;
  static constexpr uint32_t kIBytes = 2u * kKBlockBytes;
  static constexpr uint32_t kXKBlock = kMmaN * 128u;               // one K block of an X stage: 64 rows x 128 B
  static constexpr uint32_t kXStageBytes = static_cast<uint32_t>(KB) * kXKBlock;
  static constexpr int kXStages = 2;                               // one per warpgroup pair
  static constexpr int kBars = 2 + 2 * kXStages + 4;
  static constexpr size_t kSmemBytes = 1024 + kWBytes + kIBytes + kXStages * kXStageBytes + kBars * 8 + 16;
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
//   G = d * families       d CTAs per family share its pairs round-robin
//   G = d * families + r   (148 SMs, 20 families: d = 7, r = 8): round-robin with d or
//       d + 1 CTAs per family, or, with CGF_BALANCE:
//       d dedicated CTAs per family plus r FLOATERS.  The pair sequence of every
//       family is cut into S equal parts; in h of them a floater joins as member
//       d + 1 (S = families / gcd, h = r / gcd).  Floater k spends its s-th part
//       on family (k*S + s) / h, so every CTA gets ~npairs * families / G pairs
//       instead of npairs / d on the d-CTA families and npairs / (d+1) on the
//       others (-5 % on the critical path at config 2).  Every CTA visits the
//       parts in increasing order and a floater's part index equals the family's
//       part index, so all look-back dependencies point to work that is running
//       or done: no cycles.
// ---------------------------------------------------------------------------
// Measured (config 2, A/B in one run): 116 us balanced vs 105.6 us round-robin.
// The floater couples the families it visits: a family helped in part q reaches
// part q+1 earlier than the family the floater goes to next, the round-robin
// chain of the new family cannot pass the floater's pairs, and the lags add up.
// Kept (off) with its coverage test; row-disjoint shares are the next thing to try.
#ifndef CGF_BALANCE
#define CGF_BALANCE 0
#endif
struct Seg { int fam, j0, stride, count; };
struct Schedule {
  int mode;        // 0 = whole families, 1 = round-robin, 2 = balanced with floaters
  int cta, G, nfam, npairs;
  int d, S, h;     // mode 2
  __host__ __device__ __forceinline__ static int gcd(int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; }
  __host__ __device__ __forceinline__ Schedule(int cta_, int G_, int nfam_, int npairs_,
                                              bool balance = CGF_BALANCE != 0)
      : cta(cta_), G(G_), nfam(nfam_), npairs(npairs_), d(0), S(1), h(0) {
    if (G < nfam) { mode = 0; return; }
    mode = 1;
    d = G / nfam;
    const int r = G - d * nfam;
    if (balance && r != 0) {
      const int g = gcd(nfam, r);
      S = nfam / g; h = r / g;
      if (npairs >= 2 * S * (d + 1)) mode = 2;
    }
  }
  __host__ __device__ __forceinline__ int nseg() const {
    if (mode == 0) return cta < nfam ? (nfam - 1 - cta) / G + 1 : 0;
    return mode == 1 ? 1 : S;
  }
  __host__ __device__ __forceinline__ static int count_from(int j0, int end, int stride) {
    return j0 < end ? (end - j0 + stride - 1) / stride : 0;
  }
  __host__ __device__ __forceinline__ Seg get(int s) const {
    Seg sg;
    if (mode == 0) {
      sg.fam = cta + s * G; sg.j0 = 0; sg.stride = 1; sg.count = npairs;
    } else if (mode == 1) {
      sg.fam = cta % nfam;
      const int rank = cta / nfam;
      sg.stride = (G - 1 - sg.fam) / nfam + 1;
      sg.j0 = rank;
      sg.count = count_from(rank, npairs, sg.stride);
    } else {
      const int b0 = (int)((long long)s * npairs / S), b1 = (int)((long long)(s + 1) * npairs / S);
      if (cta < nfam * d) {                          // dedicated: member `rank` of its family
        sg.fam = cta % nfam;
        const int rank = cta / nfam;
        // is a floater with this family in part q?  family f owns global slices [f*h, f*h + h)
        auto members = [&](int q) {
          bool helped = false;
          for (int u = sg.fam * h; u < sg.fam * h + h; ++u) helped = helped || (u % S == q);
          return d + (helped ? 1 : 0);
        };
        // the members that get one pair more than the others in a part (its length
        // is not a multiple of the member count) rotate from part to part
        int start = 0;
        for (int q = 0; q < s; ++q) {
          const int len = (int)((long long)(q + 1) * npairs / S) - (int)((long long)q * npairs / S);
          start += len % members(q);
        }
        sg.stride = members(s);
        sg.j0 = b0 + ((rank - start) % d + d) % d;
      } else {                                       // floater: member d of family (k*S + s) / h
        const int k = cta - nfam * d;
        sg.fam = (k * S + s) / h;
        sg.stride = d + 1;
        sg.j0 = b0 + d;
      }
      sg.count = count_from(sg.j0, b1, sg.stride);
    }
    return sg;
  }
};

// ---------------------------------------------------------------------------
// The fused kernel.  KB = head width / 64 (K blocks of the gate GEMMs).
// ---------------------------------------------------------------------------
// (registers are allocated in units of four warps: 18 warps count as 20, which
// caps the kernel at 96 registers per thread)
// MUL: multiply the output by a second activation tensor on the way out (the
// gating product of RecurrentBlock), SURVEY.md section 8(f) row F2.
template <int KB, bool FAST, bool DBG, bool MUL>
__global__ void __launch_bounds__(kThreads, 1)
rglru_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const FusedParams p) {
  using Cfg = FusedCfg<KB>;
  constexpr int CBS = KB / 2;               // 128-channel halves per head
  constexpr int XS = Cfg::kXStages;
  constexpr uint32_t IDESC = umma_idesc(kMch, kMmaN);

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
  uint64_t* t_full = x_empty + XS;          // [2] accumulators of pair p are complete
  uint64_t* t_empty = t_full + 2;           // [2] both warpgroups of pair p have read them
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    mbar_init(w_empty, 1);
    for (int i = 0; i < XS; ++i) { mbar_init(x_full + i, 1); mbar_init(x_empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, 8); }
    fence_mbar_init();
  }
  if (warp == kEpiWarps) tmem_alloc(smem_u32(tmem_holder), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  // work schedule of this CTA (identical in every role), see Schedule above.
  // Ticket = tt * B + b (time-major); pair j = tickets 2j, 2j + 1 = one MMA tile.
  const int nfam = p.families;
  const int ntiles = p.ntt * p.B;
  const int npairs = (ntiles + 1) >> 1;
  const Schedule sched(blockIdx.x, gridDim.x, nfam, npairs);
  const int nsegs = sched.nseg();

  if (warp == kEpiWarps) {
    // ===================================================== TMA producer
    {
      uint32_t mq = 0, witer = 0;
      int tn = 0; (void)tn;
      int cur_fam = -1;
      for (int sgi = 0; sgi < nsegs; ++sgi) {
        const Seg sg = sched.get(sgi);
        if (sg.count == 0) continue;
        const int fam = sg.fam;
        if (fam != cur_fam) {
        cur_fam = fam;
        if (witer > 0) mbar_wait(w_empty, (witer - 1) & 1, p.err, 1);
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
          CGF_EVENT(0, 1);
          mbar_wait(x_empty + stage, (use & 1) ^ 1, p.err, 2);
          CGF_EVENT(0, 2);
          if (p.conv_flags != nullptr) {
            // x is being written right now by the Conv1D producer kernel on another
            // stream (conv1d_w4_stream_kernel): wait until every channel tile of
            // the 64-step group(s) of this MMA tile is there, then order the
            // generic-proxy view before the async-proxy (TMA) reads.
            for (int hf = 0; hf < nhalf; ++hf) {
              const int ticket = t1st + hf;
              const int tt = ticket / p.B, b = ticket - tt * p.B;
              const int* flag = p.conv_flags + (tt >> 1) * p.B + b;
              const long long t0w = clock64();
              unsigned polls = 0;
              while (ld_acquire(flag) < p.conv_need) {
                __nanosleep(100);
                polls += 15;
                if (watchdog_expired(t0w, p.err, 8, polls)) break;
              }
            }
            asm volatile("fence.proxy.async;" ::: "memory");
          }
          if (elect_one()) mbar_expect_tx(x_full + stage, nhalf * (Cfg::kXStageBytes / 2));
          for (int hf = 0; hf < nhalf; ++hf) {
            const int ticket = t1st + hf;
            const int tt = ticket / p.B, b = ticket - tt * p.B;
            if (elect_one()) {
#pragma unroll
              for (int kb = 0; kb < KB; ++kb)
                tma_load_3d(sX + stage * Cfg::kXStageBytes + kb * Cfg::kXKBlock + hf * (kTile * 128), &tmap_x,
                            x_full + stage, c_head + kb * 64, tt * kTile, b);
            }
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();
  } else if (warp == kEpiWarps + 1) {
    // ===================================================== MMA issuer
    {
      uint32_t mq = 0, witer = 0;
      int tn = 0; (void)tn;
      int cur_fam = -1;
      for (int sgi = 0; sgi < nsegs; ++sgi) {
        const Seg sg = sched.get(sgi);
        if (sg.count == 0) continue;
        const int fam = sg.fam;
        if (fam != cur_fam) {
          if (cur_fam >= 0) {                            // the old family's weights may be overwritten
            if (elect_one()) umma_commit(w_empty);
            __syncwarp();
          }
          cur_fam = fam;
          mbar_wait(w_full, witer & 1, p.err, 3);
          tc_fence_after();
          ++witer;
        }
        const int cb = fam % CBS;
#pragma unroll 1
        for (int m = 0; m < sg.count; ++m, ++mq) {
          const uint32_t pr = mq & 1u, use = mq >> 1;      // warpgroup pair == X stage
          CGF_EVENT(1, 1);
          mbar_wait(x_full + pr, use & 1, p.err, 4);
          CGF_EVENT(1, 2);
          mbar_wait(t_empty + pr, (use & 1) ^ 1, p.err, 5);
          CGF_EVENT(1, 3);
          tc_fence_after();
          const uint32_t xs = sX + pr * Cfg::kXStageBytes;
          const uint32_t dcol = tmem_base + pr * kPairCols;
          if (elect_one()) {
#pragma unroll
          for (int gate = 0; gate < 2; ++gate) {
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_bf16(dcol + gate * kMmaN, umma_desc(sW + (gate * KB + kb) * kKBlockBytes + k * 32),
                          umma_desc(xs + kb * Cfg::kXKBlock + k * 32), IDESC, (kb | k) != 0);
              }
            }
          }
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_bf16(dcol + 2 * kMmaN, umma_desc(sI + kb * kKBlockBytes + k * 32),
                        umma_desc(xs + (2 * cb + kb) * Cfg::kXKBlock + k * 32), IDESC, (kb | k) != 0);
            }
          }
          umma_commit(t_full + pr);
          umma_commit(x_empty + pr);
          }
          __syncwarp();
          CGF_EVENT(1, 4);
        }
      }
    }
    __syncwarp();
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
        constexpr int kLook = CGF_LOOK;
        int end = tt;                                    // tiles [end, tt): aggregate seen
        const long long t_start = clock64();
        unsigned polls = 0;
        c0 = 0.0f;
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
                if (__all_sync(0xffffffffu, (unsigned)wap[k] == epoch && (unsigned)wah[k] == epoch))
                  used = k + 1;
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
          const size_t src = pd.widx - (size_t)(tt - k) * wstep;
          c0 = fmaf(tagged_value(ld_relaxed_u64(p.agg_p + src)), c0,
                    tagged_value(ld_relaxed_u64(p.agg_h + src)));
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
      const int E = p.E;
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
      const int E = p.E;
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
      // CGF_LATE_WAIT_ST: the tile's state stores are awaited here, right before they
      // are read back, instead of right after they were issued
      if (CGF_LATE_WAIT_ST) tmem_wait_st();
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

    int cur_fam = -1, ch = 0;
    uint32_t bx2 = 0, ba2 = 0, sp2 = 0;
    for (int sgi = 0; sgi < nsegs; ++sgi) {
      const Seg sg = sched.get(sgi);
      if (sg.count == 0) continue;
      const int fam = sg.fam;
      if (fam != cur_fam) {                                // per-channel constants of the new family
        cur_fam = fam;
        ch = fam * kMch + chl;
        bx2 = 0; ba2 = 0;
        if (p.bias_x != nullptr) { const uint32_t v = p.bias_x[ch]; bx2 = v | (v << 16); }
        if (p.bias_a != nullptr) { const uint32_t v = p.bias_a[ch]; ba2 = v | (v << 16); }
        sp2 = p.neg8sp_bf[ch]; sp2 |= sp2 << 16;
      }
#pragma unroll 1
      for (int m = 0; m < sg.count; ++m, ++mq) {
        if ((mq & 1u) != pr) continue;
        const uint32_t use = mq >> 1;
        const int ticket = 2 * (sg.j0 + m * sg.stride) + (int)hf;
        if (twarp) CGF_EVENT(trole, 8);
        const unsigned long long early = request_pred();
        if (twarp) CGF_EVENT(trole, 1);
        // CGF_WAIT_LATE: the pending tile needs nothing from the new accumulators --
        // finish it first, so that a late MMA costs no idle time in this warpgroup
        const bool pend_first = CGF_WAIT_LATE && !CGF_INTERLEAVE && pd.on;
        if (pend_first) finish(early);
        mbar_wait(t_full + pr, use & 1, p.err, 6);
        if (twarp) CGF_EVENT(trole, 2);
        tc_fence_after();
        if (ticket >= ntiles) {                            // odd tile count: nothing in my half
          release_slot();
          if (pd.on) finish(early);
          continue;
        }
        const int tt = ticket / p.B, b = ticket - tt * p.B;
        const int t0 = tt * kTile;
        const unsigned rbits = p.reset_bits[(long long)b * p.bits_bstride + tt];   // kTile == 32: one word
        const int nvalid = p.T - t0;                       // >= 1; >= kTile for a full tile
        const bool fast_tile = CGF_PRELOAD && rbits == 0u && nvalid >= kTile;
        // F(k-1): before the new tile -- or, for a common-case tile, only the carry
        // now and the replay inside the gate loop below
        const bool weave = CGF_INTERLEAVE && !MUL && fast_tile && pd.on;
        float ph = 0.0f;                                   // running state of the pending tile's replay
        if (weave) ph = resolve_carry(early);
        else if (pd.on) finish(early);
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
            // the pending tile's (a, x~) of these 16 steps leave my state columns
            // before this tile's take their place
            uint32_t os[2][8];
            if (weave) {
              tmem_ld8(tm_state + hh * 16, os[0]);
              tmem_ld8(tm_state + hh * 16 + 8, os[1]);
              tmem_wait_ld();
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t st[8];
#pragma unroll
              for (int i = 0; i < 4; ++i)
                gate_step(xv[c * 4 + i], gx[c * 4 + i], ga[c * 4 + i], hh * 16 + c * 8 + 2 * i, FalseTag{}, st[i], st[4 + i]);
              tmem_st8(tm_state + hh * 16 + c * 8, st);
              if (weave) replay_chunk(hh * 2 + c, os[c], ph, nullptr);
            }
          }
          if (weave) retire_pending(ph);
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
        if (!CGF_LATE_WAIT_ST) tmem_wait_st();
        if (twarp) CGF_EVENT(trole, 4);
        // publish the tile's aggregate (tile 0 publishes its state right away in
        // F instead) and queue the tile for F
        pd.on = true; pd.tt = tt; pd.b = b; pd.nvalid = nvalid; pd.ch = ch; pd.P = P; pd.H = Hh;
        pd.widx = (((size_t)fam * p.ntt + tt) * p.B + b) * kMch + chl;
        pd.yp = p.y + ((size_t)b * p.T + t0) * p.E + ch;
        if (tt > 0 && tt + 1 < p.ntt) {
          st_relaxed_u64(p.agg_p + pd.widx, pack_tagged(P, epoch));
          st_relaxed_u64(p.agg_h + pd.widx, pack_tagged(Hh, epoch));
        }
        if (twarp) CGF_EVENT(trole, 7);
      }
    }
    if (pd.on) finish(request_pred());
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
