// C ABI of the B200 recurrent hot path (see include/cadence_b200.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared
//        -Xcompiler -fPIC  (cadence_gemma_b200/build.py)
// No --use_fast_math: the "exact" arithmetic modes rely on IEEE division,
// sqrt.rn and the accurate expf/log1pf of the CUDA math library.
#include "../../include/cadence_b200.h"

#include <cuda_runtime.h>
#include <cstdlib>
#include <stdint.h>

#include <mutex>

#include "cg_common.cuh"
#include "cg_conv1d.cuh"
#include "cg_scan.cuh"
#include "cg_fused.cuh"
#include "cg_decode.cuh"
#include "cg_train.cuh"

namespace {

using cg::ConvParams;
using cg::ScanParams;

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

constexpr int kMinSuperChunk = 64;  // smallest NW * 4 * L of any compiled geometry

// Per-device launch state: cudaFuncSetAttribute, occupancy and the SM count belong
// to a device, not to the process (a process may drive several GPUs).
constexpr int kMaxDevices = 64;
struct DeviceSlot { std::once_flag once; int sms = 0; int resident = 0; cudaError_t err = cudaSuccess; };

// scratch layout: [ticket, epoch | 256 B][agg_p][agg_h][pref][neg8sp]
// The exchange arrays hold 64-bit {value, epoch} words; the scratch must be
// zero-filled once before its first use and is self-cleaning afterwards.
struct Workspace {
  int* counter; unsigned* epoch;
  unsigned long long *agg_p, *agg_h, *pref;
  float* neg8sp; uint16_t* neg8sp_bf; unsigned* reset_bits;
  size_t total;
};

// EC: channels per column tile, SC: time steps per work item (super-chunk).
Workspace carve(void* base, int B, int T, int E, int EC, int SC) {
  Workspace w;
  const size_t ctiles = (E + EC - 1) / EC;
  const size_t nitems = (size_t)B * ctiles * ((T + SC - 1) / SC);
  char* p = reinterpret_cast<char*>(base);
  size_t off = 0;
  w.counter = reinterpret_cast<int*>(p + off);
  w.epoch = reinterpret_cast<unsigned*>(p + off + 4);
  off += 256;
  w.agg_p = reinterpret_cast<unsigned long long*>(p + off); off += nitems * EC * 8;
  w.agg_h = reinterpret_cast<unsigned long long*>(p + off); off += nitems * EC * 8;
  w.pref = reinterpret_cast<unsigned long long*>(p + off); off += nitems * EC * 8;
  w.neg8sp = reinterpret_cast<float*>(p + off); off += round_up((size_t)E * sizeof(float), 256);
  w.neg8sp_bf = reinterpret_cast<uint16_t*>(p + off); off += round_up((size_t)E * sizeof(uint16_t), 256);
  w.reset_bits = reinterpret_cast<unsigned*>(p + off);
  off += round_up((size_t)B * ((T + 31) / 32) * sizeof(unsigned), 256);
  w.total = off;
  return w;
}

template <typename IO, int KIND, int ARITH, int L, int NW, int STAGES, int MINB>
int launch_scan(ScanParams p, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  constexpr int EC = cg::kCvl * cg::IoVec<IO>::V;
  constexpr int SC = cg::kSegs * L * NW;
  static_assert(SC >= kMinSuperChunk, "update kMinSuperChunk");
  constexpr size_t smem = cg::scan_smem_bytes<IO, KIND, ARITH, L, NW, STAGES>();
  if (smem > 227 * 1024) return CG_ERR_MODE;   // geometry does not fit this dtype / mode
  const Workspace ws = carve(workspace, p.B, p.T, p.E, EC, SC);
  if (ws.total > workspace_bytes) return CG_ERR_WORKSPACE;
  p.ctiles = (p.E + EC - 1) / EC;
  p.ncols = p.B * p.ctiles;
  p.nchunks = (p.T + SC - 1) / SC;
  p.nitems = p.ncols * p.nchunks;
  p.counter = ws.counter; p.epoch = ws.epoch;
  p.agg_p = ws.agg_p; p.agg_h = ws.agg_h; p.pref = ws.pref;
  auto kernel = cg::scan_kernel<IO, KIND, ARITH, L, NW, STAGES, MINB>;
  // CTAs that fit on the device: function attributes and the occupancy belong to a
  // device, so they are set up once per device (a process may drive several GPUs)
  static DeviceSlot slots[kMaxDevices];
  int dev = 0;
  if (cudaError_t e = cudaGetDevice(&dev)) return (int)e;
  if (dev < 0 || dev >= kMaxDevices) return (int)cudaErrorInvalidDevice;
  DeviceSlot& slot = slots[dev];
  std::call_once(slot.once, [&] {
    slot.err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (slot.err != cudaSuccess) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    int sms = 0, per_sm = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    slot.err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, NW * 32, smem);
    if (slot.err != cudaSuccess) return;
    if (per_sm < 1 || sms < 1) { slot.err = cudaErrorLaunchOutOfResources; return; }
    slot.sms = per_sm * sms;   // resident CTAs
  });
  if (slot.err != cudaSuccess) return (int)slot.err;
  const int resident = slot.sms;
  const int grid = p.nitems < resident ? p.nitems : resident;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NW * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // behind the prologue, see scan_kernel
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, p)) return (int)e;
  return (int)cudaGetLastError();
}

template <typename IO, int KIND, int ARITH>
int launch_strict(const ScanParams& p, cudaStream_t stream) {
  dim3 grid((p.E + 127) / 128, p.B);
  cg::strict_scan_kernel<IO, KIND, ARITH><<<grid, 128, 0, stream>>>(p);
  return (int)cudaGetLastError();
}

// Kernel geometry per variant id: L steps per lane, NW warps (= chunks) per
// CTA, staging buffers, minimum resident CTAs per SM (caps registers).
// Variant 0 is the default; the others exist for the tuning scripts.
template <typename IO, int KIND, int ARITH>
int dispatch_geometry(int variant, const ScanParams& p, void* ws, size_t ws_bytes,
                      cudaStream_t stream) {
  switch (variant) {
    case 0: return launch_scan<IO, KIND, ARITH, 8, 4, 1, 4>(p, ws, ws_bytes, stream);   // default
    case 1: return launch_scan<IO, KIND, ARITH, 4, 8, 1, 3>(p, ws, ws_bytes, stream);
    case 2: return launch_scan<IO, KIND, ARITH, 4, 4, 1, 6>(p, ws, ws_bytes, stream);
    case 3: return launch_scan<IO, KIND, ARITH, 8, 2, 1, 8>(p, ws, ws_bytes, stream);
    case 4: return launch_scan<IO, KIND, ARITH, 4, 8, 2, 2>(p, ws, ws_bytes, stream);
    case 5: return launch_scan<IO, KIND, ARITH, 8, 8, 1, 2>(p, ws, ws_bytes, stream);
    default: return CG_ERR_MODE;
  }
}

int check_common(int B, int T, int E, int dtype) {
  if (B < 1 || T < 1 || E < 1) return CG_ERR_SHAPE;
  if (dtype != CG_DTYPE_F32 && dtype != CG_DTYPE_BF16) return CG_ERR_DTYPE;
  return CG_OK;
}

}  // namespace

extern "C" {

int cg_abi_version(void) { return CG_ABI_VERSION; }

const char* cg_status_string(int status) {
  switch (status) {
    case CG_OK: return "ok";
    case CG_ERR_NULL: return "null pointer argument";
    case CG_ERR_SHAPE: return "invalid shape (B, T, E, W must be >= 1; decode needs T == 1)";
    case CG_ERR_DTYPE: return "unsupported dtype (fp32 = 0, bf16 = 1)";
    case CG_ERR_ALIGN: return "pointers must be 16-byte aligned and E / row strides a multiple of 16 bytes";
    case CG_ERR_WORKSPACE: return "workspace too small (see cg_scan_workspace_bytes)";
    case CG_ERR_MODE: return "unsupported arith_mode / mask_mode / variant";
    case CG_ERR_UNSUPPORTED: return "shape / dtype not supported by the fused tensor-core path (use cg_rglru_fwd)";
    default: break;
  }
  if (status > 0) return cudaGetErrorString(static_cast<cudaError_t>(status));
  return "unknown status";
}

size_t cg_scan_workspace_bytes(int B, int T, int E, int dtype) {
  if (B < 1 || T < 1 || E < 1) return 0;
  const int V = dtype == CG_DTYPE_BF16 ? 8 : 4;
  return carve(nullptr, B, T, E, cg::kCvl * V, kMinSuperChunk).total;
}

int cg_conv1d_fwd(const void* x, const void* w, const void* b, const void* seg, int seg_is_i64,
                  long long seg_batch_stride, void* y, void* cache_out, int B, int T, int E, int W,
                  int dtype, int mask_mode, int arith_mode, cg_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!x || !w || !b || !seg || !y) return CG_ERR_NULL;
  if (int rc = check_common(B, T, E, dtype)) return rc;
  if (W < 1 || W > 17) return CG_ERR_SHAPE;
  if (mask_mode != CG_MASK_FORK && mask_mode != CG_MASK_UPSTREAM) return CG_ERR_MODE;
  const bool bf = dtype == CG_DTYPE_BF16;
  const bool emul = (arith_mode & CG_ARITH_FP32) == 0;
  ConvParams p{};
  p.x = x; p.w = w; p.bias = b; p.seg = seg; p.seg_bstride = seg_batch_stride;
  p.seg_is_i64 = seg_is_i64; p.y = y; p.cache_out = cache_out;
  p.B = B; p.T = T; p.E = E; p.W = W; p.mask_mode = mask_mode;
  const int V = bf ? 8 : 4;
  const bool vec_ok = W == 4 && E % V == 0 && aligned16(x) && aligned16(w) && aligned16(b) &&
                      aligned16(y) && (!cache_out || aligned16(cache_out));
  if (vec_ok) {
    const int variant = (arith_mode >> 8) & 0xff;
    // LC rows per thread (+3 halo rows re-read through L2): 8 by default
#define CG_CONV(LCV)                                                                       \
    {                                                                                      \
      const int tslots = (T + LCV - 1) / LCV;                                              \
      dim3 grid((E + 8 * V - 1) / (8 * V), (tslots + 15) / 16, B);                         \
      if (bf && emul) cg::conv1d_w4_kernel<uint16_t, true, LCV><<<grid, 128, 0, stream>>>(p);   \
      else if (bf) cg::conv1d_w4_kernel<uint16_t, false, LCV><<<grid, 128, 0, stream>>>(p);     \
      else cg::conv1d_w4_kernel<float, true, LCV><<<grid, 128, 0, stream>>>(p);            \
    }
    if (variant == 1) CG_CONV(16) else if (variant == 2) CG_CONV(4) else CG_CONV(8)
#undef CG_CONV
  } else {
    if (T > 65535 || B > 65535) return CG_ERR_SHAPE;
    dim3 grid((E + 127) / 128, T, B);
    if (bf && emul) cg::conv1d_generic_kernel<uint16_t, true><<<grid, 128, 0, stream>>>(p);
    else if (bf) cg::conv1d_generic_kernel<uint16_t, false><<<grid, 128, 0, stream>>>(p);
    else cg::conv1d_generic_kernel<float, true><<<grid, 128, 0, stream>>>(p);
  }
  return (int)cudaGetLastError();
}

namespace {
constexpr int kConvBwdLC = 4;
inline int conv_bwd_tblocks(int T) {      // CTAs along time: kConvBwdSpan blocks of 16 slots x LC steps each
  const int blocks = ((T + kConvBwdLC - 1) / kConvBwdLC + 15) / 16;
  return (blocks + cg::kConvBwdSpan - 1) / cg::kConvBwdSpan;
}
}  // namespace

size_t cg_conv1d_bwd_workspace_bytes(int B, int T, int E) {
  if (B < 1 || T < 1 || E < 1) return 0;
  return (size_t)B * conv_bwd_tblocks(T) * 5 * E * sizeof(float);
}

int cg_conv1d_bwd(const void* gy, const void* x, const void* w, const void* seg, int seg_is_i64,
                  long long seg_batch_stride, void* dx, void* dw, void* db, void* workspace,
                  size_t workspace_bytes, int B, int T, int E, int W, int dtype, int mask_mode,
                  cg_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!gy || !x || !w || !seg || !dx || !dw || !db || !workspace) return CG_ERR_NULL;
  if (int rc = check_common(B, T, E, dtype)) return rc;
  if (mask_mode != CG_MASK_FORK && mask_mode != CG_MASK_UPSTREAM) return CG_ERR_MODE;
  if (W != 4) return CG_ERR_UNSUPPORTED;            // the temporal width of every Griffin preset
  if (B > 65535) return CG_ERR_SHAPE;
  const bool bf = dtype == CG_DTYPE_BF16;
  const int V = bf ? 8 : 4;
  if (E % V != 0) return CG_ERR_ALIGN;
  if (!aligned16(gy) || !aligned16(x) || !aligned16(w) || !aligned16(dx) || !aligned16(workspace))
    return CG_ERR_ALIGN;
  if (cg_conv1d_bwd_workspace_bytes(B, T, E) > workspace_bytes) return CG_ERR_WORKSPACE;
  cg::ConvBwdParams p{};
  p.gy = gy; p.x = x; p.w = w; p.seg = seg; p.seg_bstride = seg_batch_stride; p.seg_is_i64 = seg_is_i64;
  p.dx = dx; p.partial = reinterpret_cast<float*>(workspace);
  p.B = B; p.T = T; p.E = E; p.mask_mode = mask_mode;
  const int tblocks = conv_bwd_tblocks(T);
  if (tblocks > 65535) return CG_ERR_SHAPE;
  dim3 grid((E + 8 * V - 1) / (8 * V), tblocks, B);
  if (bf) cg::conv1d_w4_bwd_kernel<uint16_t, kConvBwdLC><<<grid, 128, 0, stream>>>(p);
  else cg::conv1d_w4_bwd_kernel<float, kConvBwdLC><<<grid, 128, 0, stream>>>(p);
  if (cudaError_t err = cudaGetLastError()) return (int)err;
  const int nparts = B * tblocks;
  const int rgrid = (5 * E + 127) / 128;
  if (bf) cg::conv1d_bwd_reduce_kernel<uint16_t><<<rgrid, 128, 0, stream>>>(p.partial, nparts, E, dw, db);
  else cg::conv1d_bwd_reduce_kernel<float><<<rgrid, 128, 0, stream>>>(p.partial, nparts, E, dw, db);
  return (int)cudaGetLastError();
}

int cg_conv1d_decode(const void* x, const void* w, const void* b, const void* cache_in,
                     int cache_dtype, void* y, void* cache_out, int B, int E, int W, int dtype,
                     int arith_mode, cg_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!x || !w || !b || !cache_in || !y) return CG_ERR_NULL;
  if (int rc = check_common(B, 1, E, dtype)) return rc;
  if (cache_dtype != CG_DTYPE_F32 && cache_dtype != CG_DTYPE_BF16) return CG_ERR_DTYPE;
  if (W < 2 || W > 17 || B > 65535) return CG_ERR_SHAPE;
  const bool bf = dtype == CG_DTYPE_BF16;
  const bool emul = (arith_mode & CG_ARITH_FP32) == 0;
  ConvParams p{};
  p.x = x; p.w = w; p.bias = b; p.y = y; p.cache_in = cache_in; p.cache_out = cache_out;
  p.cache_is_bf16 = cache_dtype == CG_DTYPE_BF16;
  p.B = B; p.T = 1; p.E = E; p.W = W;
  dim3 grid((E + 127) / 128, B);
  if (bf && emul) cg::conv1d_decode_kernel<uint16_t, true><<<grid, 128, 0, stream>>>(p);
  else if (bf) cg::conv1d_decode_kernel<uint16_t, false><<<grid, 128, 0, stream>>>(p);
  else cg::conv1d_decode_kernel<float, true><<<grid, 128, 0, stream>>>(p);
  return (int)cudaGetLastError();
}

int cg_rglru_fwd(const void* x, const void* gemm_x, const void* gemm_a, long long gate_row_stride,
                 int gate_block_width, const void* bias_x, const void* bias_a, const void* a_param, const void* seg,
                 int seg_is_i64, long long seg_batch_stride, const float* h0, void* y,
                 float* last_h, void* workspace, size_t workspace_bytes, int B, int T, int E,
                 int dtype, int arith_mode, cg_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!x || !gemm_x || !gemm_a || !a_param || !seg || !y || !workspace) return CG_ERR_NULL;
  if (int rc = check_common(B, T, E, dtype)) return rc;
  if (B > 65535) return CG_ERR_SHAPE;
  const bool bf = dtype == CG_DTYPE_BF16;
  const int V = bf ? 8 : 4;
  const int variant = (arith_mode >> 8) & 0xff;
  const int mode = arith_mode & 7;
  const bool strict = (mode & CG_ARITH_STRICT) != 0;
  if ((arith_mode & ~0xff07) != 0) return CG_ERR_MODE;
  if (gate_row_stride < E || gate_block_width < 0) return CG_ERR_SHAPE;
  if (gate_block_width > 0 && (E % gate_block_width != 0 || gate_row_stride < 2LL * E)) return CG_ERR_SHAPE;
  if (!strict) {
    if (E % V != 0 || gate_row_stride % V != 0 || gate_block_width % V != 0) return CG_ERR_ALIGN;
    if (!aligned16(x) || !aligned16(gemm_x) || !aligned16(gemm_a) || !aligned16(y) ||
        (bias_x && !aligned16(bias_x)) || (bias_a && !aligned16(bias_a)) ||
        (h0 && !aligned16(h0)) || (last_h && !aligned16(last_h)) || !aligned16(workspace))
      return CG_ERR_ALIGN;
  }
  if (T == 1 && !strict) {   // decode step: one fused launch, no scratch traffic
    ScanParams p{};
    p.x = x; p.gemm_x = gemm_x; p.gemm_a = gemm_a; p.bias_x = bias_x; p.bias_a = bias_a;
    p.seg = seg; p.seg_bstride = seg_batch_stride; p.seg_is_i64 = seg_is_i64;
    p.gate_ld = gate_row_stride; p.gate_bw = gate_block_width; p.h0 = h0; p.y = y;
    p.last_h = last_h; p.B = B; p.T = 1; p.E = E;
    const int threads = B * (E / V);
    const int grid = (threads + 127) / 128;
    if (bf) {
      switch (mode & 3) {
        case 0: cg::rglru_step_kernel<uint16_t, 0><<<grid, 128, 0, stream>>>(p, a_param); break;
        case 1: cg::rglru_step_kernel<uint16_t, 1><<<grid, 128, 0, stream>>>(p, a_param); break;
        case 2: cg::rglru_step_kernel<uint16_t, 2><<<grid, 128, 0, stream>>>(p, a_param); break;
        default: cg::rglru_step_kernel<uint16_t, 3><<<grid, 128, 0, stream>>>(p, a_param); break;
      }
    } else if (mode & 2) {
      cg::rglru_step_kernel<float, 3><<<grid, 128, 0, stream>>>(p, a_param);
    } else {
      cg::rglru_step_kernel<float, 1><<<grid, 128, 0, stream>>>(p, a_param);
    }
    return (int)cudaGetLastError();
  }
  // ticket / epoch live at the head and -8 * softplus(a_param) at the tail of
  // the (largest-geometry) scratch layout, i.e. outside every geometry's arrays
  const Workspace ws_min = carve(workspace, B, T, E, cg::kCvl * V, kMinSuperChunk);
  if (ws_min.total > workspace_bytes) return CG_ERR_WORKSPACE;
  const int emulate = (mode & CG_ARITH_FP32) == 0;
  const int words = (T + 31) / 32;
  {
    cg::PrologueParams q{};
    q.a_param = a_param; q.neg8sp = ws_min.neg8sp; q.neg8sp_bf = ws_min.neg8sp_bf;
    q.E = E; q.is_bf16 = bf ? 1 : 0; q.emulate = emulate;
    q.counter = strict ? nullptr : ws_min.counter; q.epoch = ws_min.epoch;
    q.seg = seg; q.seg_is_i64 = seg_is_i64; q.seg_bstride = seg_batch_stride;
    q.reset_bits = strict ? nullptr : ws_min.reset_bits;
    q.rows = seg_batch_stride == 0 ? 1 : B; q.T = T; q.words_per_row = words;
    const int n = E > q.rows * words * 32 ? E : q.rows * words * 32;
    cg::scan_prologue_kernel<<<(n + 127) / 128, 128, 0, stream>>>(q);
    if (cudaError_t err = cudaGetLastError()) return (int)err;
  }

  ScanParams p{};
  p.x = x; p.gemm_x = gemm_x; p.gemm_a = gemm_a; p.bias_x = bias_x; p.bias_a = bias_a;
  p.neg8sp = ws_min.neg8sp; p.neg8sp_bf = ws_min.neg8sp_bf; p.seg = seg;
  p.reset_bits = ws_min.reset_bits; p.bits_bstride = seg_batch_stride == 0 ? 0 : words;
  p.seg_bstride = seg_batch_stride; p.seg_is_i64 = seg_is_i64;
  p.gate_ld = gate_row_stride; p.gate_bw = gate_block_width; p.h0 = h0; p.y = y; p.last_h = last_h;
  p.B = B; p.T = T; p.E = E;

  if (strict) {
    const int m = mode & 3;
    if (bf) {
      switch (m) {
        case 0: return launch_strict<uint16_t, 0, 0>(p, stream);
        case 1: return launch_strict<uint16_t, 0, 1>(p, stream);
        case 2: return launch_strict<uint16_t, 0, 2>(p, stream);
        default: return launch_strict<uint16_t, 0, 3>(p, stream);
      }
    }
    return (m & 2) ? launch_strict<float, 0, 3>(p, stream) : launch_strict<float, 0, 1>(p, stream);
  }
  if (bf) {
    switch (mode & 3) {
      case 0: return dispatch_geometry<uint16_t, 0, 0>(variant, p, workspace, workspace_bytes, stream);
      case 1: return dispatch_geometry<uint16_t, 0, 1>(variant, p, workspace, workspace_bytes, stream);
      case 2: return dispatch_geometry<uint16_t, 0, 2>(variant, p, workspace, workspace_bytes, stream);
      default: return dispatch_geometry<uint16_t, 0, 3>(variant, p, workspace, workspace_bytes, stream);
    }
  }
  return (mode & 2) ? dispatch_geometry<float, 0, 3>(variant, p, workspace, workspace_bytes, stream)
                    : dispatch_geometry<float, 0, 1>(variant, p, workspace, workspace_bytes, stream);
}

int cg_rnn_scan_fwd(const void* x, const void* a, const unsigned char* reset, const float* h0,
                    void* y, float* last_h, void* workspace, size_t workspace_bytes, int B, int T,
                    int E, int dtype, int arith_mode, cg_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!x || !a || !reset || !y) return CG_ERR_NULL;
  if (int rc = check_common(B, T, E, dtype)) return rc;
  if (B > 65535) return CG_ERR_SHAPE;
  const bool bf = dtype == CG_DTYPE_BF16;
  const int V = bf ? 8 : 4;
  const int variant = (arith_mode >> 8) & 0xff;
  if ((arith_mode & ~0xff07) != 0) return CG_ERR_MODE;
  ScanParams p{};
  p.x = x; p.a = a; p.reset = reset; p.h0 = h0; p.y = y; p.last_h = last_h;
  p.B = B; p.T = T; p.E = E;
  if (arith_mode & CG_ARITH_STRICT) {
    return bf ? launch_strict<uint16_t, 1, 0>(p, stream) : launch_strict<float, 1, 1>(p, stream);
  }
  if (!workspace) return CG_ERR_NULL;
  if (E % V != 0) return CG_ERR_ALIGN;
  if (!aligned16(x) || !aligned16(a) || !aligned16(y) || (h0 && !aligned16(h0)) ||
      (last_h && !aligned16(last_h)) || !aligned16(workspace))
    return CG_ERR_ALIGN;
  const Workspace ws_min = carve(workspace, B, T, E, cg::kCvl * V, kMinSuperChunk);
  if (ws_min.total > workspace_bytes) return CG_ERR_WORKSPACE;
  {
    const int words = (T + 31) / 32;
    cg::PrologueParams q{};
    q.counter = ws_min.counter; q.epoch = ws_min.epoch;
    q.reset = reset; q.reset_bits = ws_min.reset_bits; q.rows = B; q.T = T; q.words_per_row = words;
    cg::scan_prologue_kernel<<<(B * words * 32 + 127) / 128, 128, 0, stream>>>(q);
    if (cudaError_t err = cudaGetLastError()) return (int)err;
    p.reset_bits = ws_min.reset_bits; p.bits_bstride = words;
  }
  return bf ? dispatch_geometry<uint16_t, 1, 0>(variant, p, workspace, workspace_bytes, stream)
            : dispatch_geometry<float, 1, 1>(variant, p, workspace, workspace_bytes, stream);
}

int cg_rnn_scan_bwd(const void* gy, const float* g_last_h, const void* a, const void* h,
                    const unsigned char* reset, const float* h0, void* dx, void* da, float* dh0,
                    void* workspace, size_t workspace_bytes, int B, int T, int E, int dtype,
                    cg_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!gy || !a || !h || !reset || !dx || !da || !workspace) return CG_ERR_NULL;
  if (int rc = check_common(B, T, E, dtype)) return rc;
  if (B > 65535) return CG_ERR_SHAPE;
  const bool bf = dtype == CG_DTYPE_BF16;
  const int V = bf ? 8 : 4;
  if (E % V != 0) return CG_ERR_ALIGN;
  if (!aligned16(gy) || !aligned16(a) || !aligned16(h) || !aligned16(dx) || !aligned16(da) ||
      (h0 && !aligned16(h0)) || (g_last_h && !aligned16(g_last_h)) || (dh0 && !aligned16(dh0)) ||
      !aligned16(workspace))
    return CG_ERR_ALIGN;
  const Workspace ws_min = carve(workspace, B, T, E, cg::kCvl * V, kMinSuperChunk);
  if (ws_min.total > workspace_bytes) return CG_ERR_WORKSPACE;
  ScanParams p{};
  p.x = gy; p.a = a; p.reset = reset; p.h0 = h0; p.hprev = h; p.g_last = g_last_h;
  p.y = dx; p.da = da; p.last_h = dh0;
  p.B = B; p.T = T; p.E = E;
  {
    const int words = (T + 31) / 32;
    cg::PrologueParams q{};                     // ticket reset + new epoch (the bitmask is not used here)
    q.counter = ws_min.counter; q.epoch = ws_min.epoch;
    q.reset = reset; q.reset_bits = ws_min.reset_bits; q.rows = B; q.T = T; q.words_per_row = words;
    cg::scan_prologue_kernel<<<(B * words * 32 + 127) / 128, 128, 0, stream>>>(q);
    if (cudaError_t err = cudaGetLastError()) return (int)err;
    p.reset_bits = ws_min.reset_bits; p.bits_bstride = words;
  }
  // geometry (steps per lane, warps per CTA, stages, min CTAs / SM), see dispatch_geometry
#ifndef CG_BWD_L
#define CG_BWD_L 8
#define CG_BWD_NW 4
#define CG_BWD_ST 1
#define CG_BWD_MINB 3
#endif
#define CG_BWD_GEOM CG_BWD_L, CG_BWD_NW, CG_BWD_ST, CG_BWD_MINB
  return bf ? launch_scan<uint16_t, 2, 0, CG_BWD_GEOM>(p, workspace, workspace_bytes, stream)
            : launch_scan<float, 2, 1, CG_BWD_GEOM>(p, workspace, workspace_bytes, stream);
}

int cg_rglru_gates_fwd(const void* x, const void* pre_x, const void* pre_a, const void* a_param,
                       const unsigned char* reset, void* a, void* nx, int B, int T, int E, int dtype,
                       cg_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!x || !pre_x || !pre_a || !a_param || !reset || !a || !nx) return CG_ERR_NULL;
  if (int rc = check_common(B, T, E, dtype)) return rc;
  const bool bf = dtype == CG_DTYPE_BF16;
  const int V = bf ? 8 : 4;
  if (E % V != 0) return CG_ERR_ALIGN;
  if (!aligned16(x) || !aligned16(pre_x) || !aligned16(pre_a) || !aligned16(a) || !aligned16(nx) ||
      !aligned16(a_param))
    return CG_ERR_ALIGN;
  cg::GateParams p{};
  p.x = x; p.pre_x = pre_x; p.pre_a = pre_a; p.a_param = a_param; p.reset = reset; p.a = a; p.nx = nx;
  p.N = (long long)B * T; p.E = E;
  const long long threads = p.N * (E / V);
  const unsigned grid = (unsigned)((threads + 255) / 256);
  if (bf) cg::rglru_gates_fwd_kernel<uint16_t><<<grid, 256, 0, stream>>>(p);
  else cg::rglru_gates_fwd_kernel<float><<<grid, 256, 0, stream>>>(p);
  return (int)cudaGetLastError();
}

size_t cg_rglru_gates_bwd_workspace_bytes(int B, int T, int E) {
  if (B < 1 || T < 1 || E < 1) return 0;
  const long long n = (long long)B * T;
  return (size_t)((n + cg::kTrainRows - 1) / cg::kTrainRows) * E * sizeof(float);
}

int cg_rglru_gates_bwd(const void* x, const void* pre_x, const void* pre_a, const void* a_param,
                       const unsigned char* reset, const void* d_nx, const void* d_a, void* dx,
                       void* d_pre_x, void* d_pre_a, void* d_a_param, void* workspace,
                       size_t workspace_bytes, int B, int T, int E, int dtype, cg_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!x || !pre_x || !pre_a || !a_param || !reset || !d_nx || !d_a || !dx || !d_pre_x || !d_pre_a ||
      !d_a_param || !workspace)
    return CG_ERR_NULL;
  if (int rc = check_common(B, T, E, dtype)) return rc;
  const bool bf = dtype == CG_DTYPE_BF16;
  const int V = bf ? 8 : 4;
  if (E % V != 0) return CG_ERR_ALIGN;
  if (!aligned16(x) || !aligned16(pre_x) || !aligned16(pre_a) || !aligned16(d_nx) || !aligned16(d_a) ||
      !aligned16(dx) || !aligned16(d_pre_x) || !aligned16(d_pre_a) || !aligned16(workspace))
    return CG_ERR_ALIGN;
  if (cg_rglru_gates_bwd_workspace_bytes(B, T, E) > workspace_bytes) return CG_ERR_WORKSPACE;
  cg::GateParams p{};
  p.x = x; p.pre_x = pre_x; p.pre_a = pre_a; p.a_param = a_param; p.reset = reset;
  p.d_nx = d_nx; p.d_a = d_a; p.dx = dx; p.d_pre_x = d_pre_x; p.d_pre_a = d_pre_a;
  p.partial = reinterpret_cast<float*>(workspace);
  p.N = (long long)B * T; p.E = E;
  const long long nparts = (p.N + cg::kTrainRows - 1) / cg::kTrainRows;
  if (nparts > 0x7fffffffLL) return CG_ERR_SHAPE;
  dim3 grid((E + 8 * V - 1) / (8 * V), 1, 1);
  // blockIdx.y is limited to 65535: large N goes through several launches of <= 65535 row groups
  for (long long first = 0; first < nparts; first += 65535) {
    const long long cnt = nparts - first < 65535 ? nparts - first : 65535;
    cg::GateParams q = p;
    const size_t roff = (size_t)first * cg::kTrainRows;
    auto adv = [&](const void* ptr) -> const void* {
      return reinterpret_cast<const char*>(ptr) + roff * E * (bf ? 2 : 4);
    };
    q.x = adv(p.x); q.pre_x = adv(p.pre_x); q.pre_a = adv(p.pre_a); q.d_nx = adv(p.d_nx); q.d_a = adv(p.d_a);
    q.dx = const_cast<void*>(adv(p.dx)); q.d_pre_x = const_cast<void*>(adv(p.d_pre_x));
    q.d_pre_a = const_cast<void*>(adv(p.d_pre_a));
    q.reset = p.reset + roff; q.partial = p.partial + (size_t)first * E; q.N = p.N - (long long)roff;
    grid.y = (unsigned)cnt;
    if (bf) cg::rglru_gates_bwd_kernel<uint16_t><<<grid, 128, 0, stream>>>(q);
    else cg::rglru_gates_bwd_kernel<float><<<grid, 128, 0, stream>>>(q);
    if (cudaError_t err = cudaGetLastError()) return (int)err;
  }
  const int rgrid = (E + 127) / 128;
  if (bf) cg::column_reduce_kernel<uint16_t><<<rgrid, 128, 0, stream>>>(p.partial, (int)nparts, E, d_a_param);
  else cg::column_reduce_kernel<float><<<rgrid, 128, 0, stream>>>(p.partial, (int)nparts, E, d_a_param);
  return (int)cudaGetLastError();
}

}  // extern "C"

// ---------------------------------------------------------------------------
// Fused tensor-core path (cg_fused.cuh)
// ---------------------------------------------------------------------------
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup, so
// the library needs no link-time dependency on libcuda.
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

bool fused_shape_ok(int E, int H, int dtype) {
  if (dtype != CG_DTYPE_BF16 || H < 1 || E < 1 || E % H != 0) return false;
  const int bw = E / H;
  return bw == 128 || bw == 256;
}

// CG_B200_STEAL=0 switches the tail stealing of the fused kernels' schedule off (A/B runs)
// (CG_B200_STEAL=n, n >= 2: steal exactly n pairs per poor unit -- tuning runs)
int fused_steal_override() {
  static const int v = [] { const char* e = getenv("CG_B200_STEAL"); return e ? atoi(e) : 1; }();
  return v;
}
bool fused_steal_enabled() { return fused_steal_override() != 0; }

// MMA tiles per cluster from which the one-launch kernel hands tiles out through ticket counters
// (profiles/r3_dynamic_tickets.txt: 52-69 tiles per cluster: static 0-1 % better; 277: tickets 1.7 % better)
constexpr long long kDynamicMinTilesPerCluster = 128;
// ... and below which (at B >= 8) the plain kernel instance beats the UNI one (r3_ab_uniform_warp_index.txt)
constexpr long long kPlainInstanceMaxTilesPerCluster = 48;

// tiles per work group (cluster / CTA) of a full-device grid: the schedule decision of launch_fused, available
// before the kernel instance is chosen
long long fused_tiles_per_group(const cg::fused::FusedParams& p, int CL, bool conv) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1)
    return 0;
  const int units = conv ? p.families / CL : p.families;
  const int G = sms / CL;
  return (long long)((p.ntt * p.B + 1) / 2) * units / (G > 0 ? G : 1);
}

template <int KB, bool FAST, bool DBG, bool MUL, bool CONV, int LOOK = 2, int EFIX = 0, bool UNI = false>
int launch_fused(const CUtensorMap& tmap, const cg::fused::FusedParams& p, int grid_limit,
                 cudaStream_t stream) {
  using Cfg = cg::fused::FusedCfg<KB>;
  auto kernel = cg::fused::rglru_fused_kernel<KB, FAST, DBG, MUL, CONV, LOOK, EFIX, UNI>;
  static DeviceSlot slots[kMaxDevices];
  int dev = 0;
  if (cudaError_t e = cudaGetDevice(&dev)) return (int)e;
  if (dev < 0 || dev >= kMaxDevices) return (int)cudaErrorInvalidDevice;
  DeviceSlot& slot = slots[dev];
  std::call_once(slot.once, [&] {
    slot.err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes);
    if (slot.err == cudaSuccess) slot.err = cudaDeviceGetAttribute(&slot.sms, cudaDevAttrMultiProcessorCount, dev);
  });
  if (slot.err != cudaSuccess) return (int)slot.err;
  const int sms = slot.sms;
  if (sms < 1) return (int)cudaErrorLaunchOutOfResources;
  // one persistent CTA per SM; grid_limit > 0 (test hook) runs with fewer CTAs than
  // column families, which exercises the family loop (weight reload) of a CTA
  // CONV kernels at head width 256: the two CTAs of a head form a thread-block cluster
  constexpr int CL = CONV ? KB / 2 : 1;
  int grid = grid_limit > 0 && grid_limit < sms ? grid_limit : sms;
  grid = grid / CL * CL;
  if (grid < CL) grid = CL;
  // programmatic dependent launch behind the prologue kernel (see the
  // griddepcontrol.wait in the epilogue warps): set-up and operand loads start
  // while the prologue is still running
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(cg::fused::kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (CL > 1) {
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = CL; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
    // every CTA waits on others (look-back): all clusters of the grid must be co-resident
    if (slot.resident == 0) {
      int ncl = 0;
      if (cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, kernel, &cfg)) return (int)e;
      slot.resident = ncl > 0 ? ncl : -1;
    }
    if (slot.resident < 1) return (int)cudaErrorLaunchOutOfResources;
    if (grid > slot.resident * CL) { grid = slot.resident * CL; cfg.gridDim = dim3(grid); }
  }
  // Tail stealing (cg::fused::Schedule): G = d * units + r work groups; the d + 1 CTAs of a rich unit end
  // after npairs / (d + 1) tiles, reload weights and refill their pipeline (~2.3 tile times, measured) and
  // take x * npoor / nrich tiles of the poor units' tails; a poor unit's d CTAs end after (npairs - x) / d.
  // Equal ends:  x = (npairs / (d (d + 1)) - 2.3) / (npoor / nrich + 1 / d).  The stolen tiles are the END of
  // a time line: a helper cannot finish them before the unit's own CTAs have produced the state they start
  // from, so more than ~1.5 tiles per helper only queue up behind that point (measured: B=16, T=8192 with
  // 18 tiles per helper 882 us, with 2 863 us, without stealing 869 us; config 2: 124.3 -> 121.8 us,
  // profiles/r3_tail_stealing.txt) -- the real fix for the 7 % imbalance is a dynamic ticket per unit
  // (DESIGN.md section 9).
  cg::fused::FusedParams q = p;
  {
    const int units = CONV ? p.families / CL : p.families;
    const int G = grid / CL, d = G / units, r = G % units;
    const int ntiles = p.ntt * p.B, npairs = (ntiles + 1) / 2;
    q.steal = 0;
    if (d >= 1 && r > 0 && r * (d + 1) >= units - r && fused_steal_enabled()) {
      const double npoor = units - r, nrich = (double)r * (d + 1);
      double x = ((double)npairs / ((double)d * (d + 1)) - 2.3) / (npoor / nrich + 1.0 / d);
      const double cap = 1.5 * nrich / npoor;
      if (x > cap) x = cap;
      if (x >= 1.0) q.steal = (int)(x + 0.5);
      if (fused_steal_override() >= 2) q.steal = fused_steal_override();
      if (q.steal > npairs / 4) q.steal = npairs / 4;
    }
  }
  // CONV kernels: long kernels hand the last 10 % of every head's tiles out through ticket counters (no
  // cluster idles while tiles are left: B=16, T=8192 867 -> 831 us); short ones keep the static schedule
  // (dynamic tickets scramble the order in which the tiles of a chain are processed, the look-back waits
  // longer: config 2 121.8 static vs 126.0 us).  CG_B200_DYNAMIC=<pct static> forces either (0 = static).
  q.static_pct = 0;
  if (CONV) {
    const int units = p.families / CL, G = grid / CL;
    const long long per_cluster = (long long)((p.ntt * p.B + 1) / 2) * units / (G > 0 ? G : 1);
    if (G >= units && per_cluster >= kDynamicMinTilesPerCluster) q.static_pct = 90;
    static const int forced = [] { const char* e = getenv("CG_B200_DYNAMIC"); return e ? atoi(e) : -1; }();
    if (forced >= 0 && forced <= 100 && G >= units) q.static_pct = forced;
    if (q.static_pct > 0) q.steal = 0;
    static const int fwd = [] { const char* e = getenv("CG_B200_FORWARD"); return e ? atoi(e) : -1; }();
    // (the UNI instances are 5 % slower when both CTAs of a cluster derive a static schedule on their own
    // than when the peer follows the leader's descriptors: profiles/r3_ab_uniform_warp_index.txt)
    q.forward = fwd >= 0 ? (q.static_pct != 0 || fwd == 1) : (q.static_pct != 0 || UNI);
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, tmap, q);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaGetLastError();
}

template <int KB, bool CONV>
int dispatch_fused(bool fast, bool dbg, bool mul, const CUtensorMap& tmap,
                   const cg::fused::FusedParams& p, int grid_limit, cudaStream_t stream) {
  if (dbg) return fast ? launch_fused<KB, true, true, false, CONV>(tmap, p, grid_limit, stream)
                       : launch_fused<KB, false, true, false, CONV>(tmap, p, grid_limit, stream);
  if (mul) return fast ? launch_fused<KB, true, false, true, CONV>(tmap, p, grid_limit, stream)
                       : launch_fused<KB, false, false, true, CONV>(tmap, p, grid_limit, stream);
  // look-back depth 2 where the carry chain is the bound (small batches), 1 elsewhere (cg_fused.cuh)
  // RecurrentGemma-2B's / 9B's width with the row pitch as a compile-time constant (the shipped arithmetic
  // only): immediate offsets instead of 64-bit address arithmetic per row in the y stores and the halo loads
  // (config 2 one launch 121.7 -> 118.0 us, B=16 T=8192 856 -> 837 us; profiles/r3_ab_efix.txt)
  if (CGF_EFIX && KB == 4 && fast && (p.E == 2560 || p.E == 4096)) {
    // one-launch kernel: warp-uniform role branches (UNI) except for static schedules at B >= 8 (see the kernel)
    static const int uni_forced = [] { const char* e = getenv("CG_B200_UNI"); return e ? atoi(e) : -1; }();
    const bool uni = CONV && grid_limit == 0 &&
                     (uni_forced >= 0 ? uni_forced != 0
                                      : !(fused_tiles_per_group(p, CONV ? KB / 2 : 1, CONV) < kPlainInstanceMaxTilesPerCluster &&
                                          p.B >= 8));
    const bool deep = p.B <= CGF_LOOK_SPLIT;
#define CG_EFIX_CASE(EV)                                                                                     \
    if (p.E == EV) {                                                                                           \
      if (uni) return deep ? launch_fused<KB, true, false, false, CONV, 2, EV, CONV>(tmap, p, grid_limit, stream) \
                           : launch_fused<KB, true, false, false, CONV, 1, EV, CONV>(tmap, p, grid_limit, stream); \
      return deep ? launch_fused<KB, true, false, false, CONV, 2, EV, false>(tmap, p, grid_limit, stream)        \
                  : launch_fused<KB, true, false, false, CONV, 1, EV, false>(tmap, p, grid_limit, stream);       \
    }
    CG_EFIX_CASE(2560)
    CG_EFIX_CASE(4096)
#undef CG_EFIX_CASE
  }
  if (p.B > CGF_LOOK_SPLIT) return fast ? launch_fused<KB, true, false, false, CONV, 1>(tmap, p, grid_limit, stream)
                           : launch_fused<KB, false, false, false, CONV, 1>(tmap, p, grid_limit, stream);
  return fast ? launch_fused<KB, true, false, false, CONV>(tmap, p, grid_limit, stream)
              : launch_fused<KB, false, false, false, CONV>(tmap, p, grid_limit, stream);
}

// Shared body of cg_rglru_fused_fwd (x = Conv1D output) and cg_recurrent_prefill_fwd
// (x = Conv1D INPUT, convolution inside the kernel when conv_w != nullptr).
int fused_forward(const void* x, const void* conv_w, const void* conv_b, void* conv_cache_out, int mask_mode,
                  const void* wpack, const void* bias_x, const void* bias_a, const void* a_param,
                  const void* seg, int seg_is_i64, long long seg_batch_stride, const float* h0, void* y,
                  float* last_h, void* workspace, size_t workspace_bytes, int B, int T, int E, int H,
                  int dtype, int arith_mode, const void* gate_mul, void* debug_out, cudaStream_t stream) {
  if (!x || !wpack || !a_param || !seg || !y || !workspace) return CG_ERR_NULL;
  if (int rc = check_common(B, T, E, dtype)) return rc;
  if (!fused_shape_ok(E, H, dtype)) return CG_ERR_UNSUPPORTED;
  const int mode = arith_mode & 7;
  const int variant = (arith_mode >> 8) & 0xff;
  if ((arith_mode & ~0xff07) != 0 || (mode & (CG_ARITH_FP32 | CG_ARITH_STRICT)) != 0) return CG_ERR_MODE;
  if (variant > 64) return CG_ERR_MODE;   // variant v > 0: at most v CTAs (test hook, see launch_fused)
  if (!aligned16(x) || !aligned16(wpack) || !aligned16(y) || !aligned16(workspace)) return CG_ERR_ALIGN;
  const bool conv = conv_w != nullptr;
  if (conv) {
    if (!conv_b) return CG_ERR_NULL;
    if (mask_mode != CG_MASK_FORK && mask_mode != CG_MASK_UPSTREAM) return CG_ERR_MODE;
    if (!aligned16(conv_w) || !aligned16(conv_b) || (conv_cache_out && !aligned16(conv_cache_out))) return CG_ERR_ALIGN;
  }
  if (B > 65535) return CG_ERR_SHAPE;
  if (gate_mul != nullptr && debug_out != nullptr) return CG_ERR_MODE;   // the debug build has no product path
  const int bw = E / H;
  const int tile_t = cg::fused::kTile;
  const Workspace ws = carve(workspace, B, T, E, cg::fused::kMch, tile_t);
  if (ws.total > workspace_bytes) return CG_ERR_WORKSPACE;
  const int words = (T + 31) / 32;
  {   // epoch bump, -8*softplus(a_param), reset bitmask (shared with the unfused path)
    cg::PrologueParams q{};
    q.a_param = a_param; q.neg8sp = ws.neg8sp; q.neg8sp_bf = ws.neg8sp_bf;
    q.E = E; q.is_bf16 = 1; q.emulate = 1;
    q.counter = ws.counter; q.epoch = ws.epoch;
    q.seg = seg; q.seg_is_i64 = seg_is_i64; q.seg_bstride = seg_batch_stride;
    q.reset_bits = ws.reset_bits;
    q.rows = seg_batch_stride == 0 ? 1 : B; q.T = T; q.words_per_row = words;
    const int n = E > q.rows * words * 32 ? E : q.rows * words * 32;
    // programmatic dependent of whatever precedes it on the stream (the kernel that
    // produces x): it runs under that kernel's tail if that kernel triggers early;
    // it reads nothing that kernel writes and does not complete before it
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((n + 127) / 128);
    cfg.blockDim = dim3(128);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaError_t err = cudaLaunchKernelEx(&cfg, cg::scan_prologue_kernel, q)) return (int)err;
    if (cudaError_t err = cudaGetLastError()) return (int)err;
  }
  EncodeTiledFn encode = encode_tiled_fn();
  if (encode == nullptr) return (int)cudaErrorNotSupported;
  CUtensorMap tmap;
  {
    const cuuint64_t dims[3] = {(cuuint64_t)E, (cuuint64_t)T, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)E * 2, (cuuint64_t)T * E * 2};
    const cuuint32_t box[3] = {64, (cuuint32_t)tile_t, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(x), dims, strides, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return (int)cudaErrorInvalidValue;
  }
  cg::fused::FusedParams p{};
  const size_t wbytes = cg_rglru_gate_pack_bytes(E, H) - 2 * cg::fused::kKBlockBytes;
  p.wpack = reinterpret_cast<const unsigned char*>(wpack);
  p.ident = p.wpack + wbytes;
  p.bias_x = reinterpret_cast<const uint16_t*>(bias_x);
  p.bias_a = reinterpret_cast<const uint16_t*>(bias_a);
  p.neg8sp_bf = ws.neg8sp_bf;
  p.reset_bits = ws.reset_bits;
  p.bits_bstride = seg_batch_stride == 0 ? 0 : words;
  p.words = words;
  p.h0 = h0; p.y = reinterpret_cast<uint16_t*>(y); p.last_h = last_h;
  p.epoch = ws.epoch; p.agg_p = ws.agg_p; p.agg_h = ws.agg_h; p.pref = ws.pref;
  p.gate_mul = reinterpret_cast<const uint16_t*>(gate_mul);
  p.x_lin = reinterpret_cast<const uint16_t*>(x);
  p.conv_w = reinterpret_cast<const uint16_t*>(conv_w);
  p.conv_b = reinterpret_cast<const uint16_t*>(conv_b);
  p.conv_cache = reinterpret_cast<uint16_t*>(conv_cache_out);
  p.mask_mode = mask_mode;
  p.dbg = reinterpret_cast<uint16_t*>(debug_out);
  p.err = ws.counter + 2;   // third word of the scratch header
  p.B = B; p.T = T; p.E = E;
  p.ntt = (T + tile_t - 1) / tile_t;
  p.families = E / cg::fused::kMch;
  cg::fused::make_bdiv((uint32_t)B, p.bdiv_m, p.bdiv_s1, p.bdiv_s2);
  p.steal = -1;   // launch_fused: depends on the grid
  p.tickets = ws.counter + 8;   // words 8 .. 63 of the scratch header, zeroed by the prologue kernel
  if (conv && H > 32) return CG_ERR_UNSUPPORTED;   // one lane per head in the ticket draw
  const bool fast = (mode & CG_ARITH_FAST) != 0;
  const bool dbg = debug_out != nullptr;
  const bool mul = gate_mul != nullptr;
  if (conv)
    return bw == 256 ? dispatch_fused<4, true>(fast, dbg, mul, tmap, p, variant, stream)
                     : dispatch_fused<2, true>(fast, dbg, mul, tmap, p, variant, stream);
  return bw == 256 ? dispatch_fused<4, false>(fast, dbg, mul, tmap, p, variant, stream)
                   : dispatch_fused<2, false>(fast, dbg, mul, tmap, p, variant, stream);
}

}  // namespace

extern "C" {

int cg_rglru_fused_supported(int E, int H, int dtype) { return fused_shape_ok(E, H, dtype) ? 1 : 0; }

size_t cg_rglru_gate_pack_bytes(int E, int H) {
  if (!fused_shape_ok(E, H, CG_DTYPE_BF16)) return 0;
  const size_t bw = (size_t)E / H;
  // [E/128 families][2 gates][bw/64 K blocks][128 rows][128 B] + identity [2][128][128 B]
  return (size_t)(E / 128) * 2 * (bw / 64) * cg::fused::kKBlockBytes + 2 * cg::fused::kKBlockBytes;
}

int cg_rglru_pack_gate_weights(const void* wx, const void* wa, void* wpack, int E, int H, int dtype,
                               cg_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!wx || !wa || !wpack) return CG_ERR_NULL;
  if (!fused_shape_ok(E, H, dtype)) return CG_ERR_UNSUPPORTED;
  if (!aligned16(wpack)) return CG_ERR_ALIGN;
  const int bw = E / H;
  const size_t wbytes = cg_rglru_gate_pack_bytes(E, H) - 2 * cg::fused::kKBlockBytes;
  const long long chunks = (long long)(wbytes / 16) + 2 * 128 * 8;
  cg::fused::pack_gate_weights_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const uint16_t*>(wx), reinterpret_cast<const uint16_t*>(wa),
      reinterpret_cast<unsigned char*>(wpack), reinterpret_cast<unsigned char*>(wpack) + wbytes, H, bw);
  return (int)cudaGetLastError();
}

size_t cg_rglru_fused_workspace_bytes(int B, int T, int E) {
  if (B < 1 || T < 1 || E < 1) return 0;
  return carve(nullptr, B, T, E, cg::fused::kMch, cg::fused::kTile).total;
}

int cg_rglru_fused_fwd(const void* x, const void* wpack, const void* bias_x, const void* bias_a,
                       const void* a_param, const void* seg, int seg_is_i64, long long seg_batch_stride,
                       const float* h0, void* y, float* last_h, void* workspace, size_t workspace_bytes,
                       int B, int T, int E, int H, int dtype, int arith_mode, const void* gate_mul,
                       cg_stream_t stream_) {
  return fused_forward(x, nullptr, nullptr, nullptr, CG_MASK_FORK, wpack, bias_x, bias_a, a_param, seg,
                       seg_is_i64, seg_batch_stride, h0, y, last_h, workspace, workspace_bytes, B, T, E, H,
                       dtype, arith_mode, gate_mul, nullptr, reinterpret_cast<cudaStream_t>(stream_));
}

int cg_recurrent_prefill_fwd(const void* x, const void* conv_w, const void* conv_b, const void* wpack,
                             const void* bias_x, const void* bias_a, const void* a_param, const void* seg,
                             int seg_is_i64, long long seg_batch_stride, const float* h0, const void* gate_mul,
                             void* y, void* conv_cache_out, float* last_h, void* workspace,
                             size_t workspace_bytes, int B, int T, int E, int H, int W, int dtype,
                             int mask_mode, int arith_mode, cg_stream_t stream_) {
  if (!conv_w || !conv_b) return CG_ERR_NULL;
  if (W != 4) return CG_ERR_UNSUPPORTED;   // the temporal width of every Griffin preset
  return fused_forward(x, conv_w, conv_b, conv_cache_out, mask_mode, wpack, bias_x, bias_a, a_param, seg,
                       seg_is_i64, seg_batch_stride, h0, y, last_h, workspace, workspace_bytes, B, T, E, H,
                       dtype, arith_mode, gate_mul, nullptr, reinterpret_cast<cudaStream_t>(stream_));
}

/* experimental (include/cadence_b200_experimental.h): the same two entry points with a
 * debug tap -- debug_out [3][B,T,E] receives the rounded gate pre-activations and the
 * (convolved) x as the epilogue sees them.  Not part of the drop-in surface. */
int cg_recurrent_prefill_debug(const void* x, const void* conv_w, const void* conv_b, const void* wpack,
                               const void* bias_x, const void* bias_a, const void* a_param, const void* seg,
                               int seg_is_i64, long long seg_batch_stride, const float* h0, void* y,
                               void* conv_cache_out, float* last_h, void* workspace, size_t workspace_bytes,
                               int B, int T, int E, int H, int dtype, int mask_mode, int arith_mode,
                               void* debug_out, cg_stream_t stream_) {
  if (!debug_out) return CG_ERR_NULL;
  return fused_forward(x, conv_w, conv_b, conv_cache_out, mask_mode, wpack, bias_x, bias_a, a_param, seg,
                       seg_is_i64, seg_batch_stride, h0, y, last_h, workspace, workspace_bytes, B, T, E, H,
                       dtype, arith_mode, nullptr, debug_out, reinterpret_cast<cudaStream_t>(stream_));
}


int cg_recurrent_decode_supported(int E, int H, int W, int dtype) {
  if (dtype != CG_DTYPE_BF16 || W != 4 || H < 1 || E < 1 || E % H != 0) return 0;
  const int bw = E / H;
  return (bw % 64 == 0 && bw <= cg::kDecMaxBw) ? 1 : 0;
}

int cg_recurrent_decode_step(const void* x, const void* conv_w, const void* conv_b, const void* cache_in,
                             int cache_dtype, const void* wx, const void* wa, const void* bias_x,
                             const void* bias_a, const void* a_param, const void* seg, int seg_is_i64,
                             long long seg_batch_stride, const float* h0, const void* gate_mul, void* y,
                             void* cache_out, float* last_h, int B, int E, int H, int W, int dtype,
                             int arith_mode, cg_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!x || !conv_w || !conv_b || !cache_in || !wx || !wa || !a_param || !seg || !y) return CG_ERR_NULL;
  if (int rc = check_common(B, 1, E, dtype)) return rc;
  if (cache_dtype != CG_DTYPE_F32 && cache_dtype != CG_DTYPE_BF16) return CG_ERR_DTYPE;
  if (!cg_recurrent_decode_supported(E, H, W, dtype)) return CG_ERR_UNSUPPORTED;
  if ((arith_mode & ~CG_ARITH_FAST) != 0) return CG_ERR_MODE;       // reference rounding, exact or fast math
  if (cache_out == cache_in) return CG_ERR_MODE;                    // several CTAs read the cache rows
  if ((B + cg::kDecBT - 1) / cg::kDecBT > 65535) return CG_ERR_SHAPE;
  cg::DecodeParams p{};
  p.x = reinterpret_cast<const uint16_t*>(x);
  p.conv_w = reinterpret_cast<const uint16_t*>(conv_w);
  p.conv_b = reinterpret_cast<const uint16_t*>(conv_b);
  p.cache_in = cache_in; p.cache_is_bf16 = cache_dtype == CG_DTYPE_BF16;
  p.wx = reinterpret_cast<const uint16_t*>(wx); p.wa = reinterpret_cast<const uint16_t*>(wa);
  p.bias_x = reinterpret_cast<const uint16_t*>(bias_x); p.bias_a = reinterpret_cast<const uint16_t*>(bias_a);
  p.a_param = reinterpret_cast<const uint16_t*>(a_param);
  p.seg = seg; p.seg_is_i64 = seg_is_i64; p.seg_bstride = seg_batch_stride;
  p.h0 = h0; p.gate_mul = reinterpret_cast<const uint16_t*>(gate_mul);
  p.y = reinterpret_cast<uint16_t*>(y); p.cache_out = cache_out; p.last_h = last_h;
  p.B = B; p.E = E; p.H = H; p.bw = E / H;
  dim3 grid(H * (p.bw / 64), (B + cg::kDecBT - 1) / cg::kDecBT);
  size_t smem = (size_t)2 * p.bw * 128;                  // both gates' [bw x 64] bf16 slice ...
  if (smem < 32768) smem = 32768;                        // ... later reused for the partial sums
  static DeviceSlot slots[kMaxDevices];
  int dev = 0;
  if (cudaError_t e = cudaGetDevice(&dev)) return (int)e;
  if (dev < 0 || dev >= kMaxDevices) return (int)cudaErrorInvalidDevice;
  std::call_once(slots[dev].once, [&] {
    slots[dev].err = cudaFuncSetAttribute(cg::recurrent_decode_kernel<true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * cg::kDecMaxBw * 128);
    if (slots[dev].err == cudaSuccess)
      slots[dev].err = cudaFuncSetAttribute(cg::recurrent_decode_kernel<false>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * cg::kDecMaxBw * 128);
  });
  if (slots[dev].err != cudaSuccess) return (int)slots[dev].err;
  if (arith_mode & CG_ARITH_FAST) cg::recurrent_decode_kernel<true><<<grid, 256, smem, stream>>>(p);
  else cg::recurrent_decode_kernel<false><<<grid, 256, smem, stream>>>(p);
  return (int)cudaGetLastError();
}

}  // extern "C"
