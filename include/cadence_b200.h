/*
 * cadence_b200.h -- C ABI of the B200 (sm_100a) recurrent hot path.
 *
 * Drop-in boundary for the recurrent block of surakku/cadence-gemma
 * (reference paths relative to the reference checkout):
 *
 *   cg_conv1d_fwd / cg_conv1d_decode  replace the body of
 *       Conv1D.forward           recurrentgemma/torch/layers.py:458-546
 *   cg_rglru_fwd                 replaces everything in RGLRU.forward after the
 *       two BlockDiagonalLinear GEMMs (gate math + rnn_scan)
 *                                recurrentgemma/torch/layers.py:345-375
 *   cg_rnn_scan_fwd              replaces rnn_scan
 *                                recurrentgemma/torch/layers.py:146-199
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     stated otherwise; tensors are dense row-major [B,T,E] with E contiguous;
 *   - the callee allocates nothing, frees nothing and keeps no state between
 *     calls; outputs and scratch are caller-provided;
 *   - all work is enqueued on `stream` (a cudaStream_t); no host<->device
 *     synchronisation, no host reads of device memory: calls are CUDA-graph
 *     capturable;
 *   - inputs are read-only (the reference's in-place masking of Conv1D's input,
 *     layers.py:524, is intentionally not reproduced);
 *   - return value: 0 = success, < 0 = argument error (cg_status_string),
 *     > 0 = a cudaError_t raised by a launch.  Nothing is thrown across the ABI.
 */
#ifndef CADENCE_B200_H_
#define CADENCE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 2: cg_rglru_fused_fwd lost its conv_flags / debug_out parameters (round-1 experiments),
 *    cg_recurrent_prefill_fwd added, cg_conv1d_stream_* and cg_rglru_fused_schedule removed.
 * Bumped on EVERY signature change; binders must compare it with cg_abi_version(). */
#define CG_ABI_VERSION 2

/* activation dtype of x / y / gate GEMM outputs / parameters */
#define CG_DTYPE_F32 0
#define CG_DTYPE_BF16 1

/* arithmetic modes (bit field)
 *   bit 0  keep fp32 in registers between load and the final store instead of
 *          reproducing the reference's per-op rounding to the tensor dtype
 *   bit 1  fast transcendentals (ex2.approx / rcp.approx / sqrt.approx, one
 *          shared reciprocal for both sigmoids, a^2 = a*a)
 *   bit 2  strict sequential kernel (one thread per channel, no FMA
 *          contraction): the device-side oracle, not a performance path
 *   bits 8..15  kernel geometry variant (0 = default); used by the tuning
 *          scripts only, results are identical across variants of one mode
 * For fp32 tensors bit 0 is irrelevant (every op already rounds to fp32).
 */
#define CG_ARITH_REFERENCE 0
#define CG_ARITH_FP32 1
#define CG_ARITH_FAST 2
#define CG_ARITH_STRICT 4
#define CG_ARITH_VARIANT(v) (((v) & 0xff) << 8)

/* Conv1D document-mask variants */
#define CG_MASK_FORK 0     /* layers.py:629-632 as the fork has it (default) */
#define CG_MASK_UPSTREAM 1 /* the commented-out upstream loop, layers.py:620-627 */

/* status codes */
#define CG_OK 0
#define CG_ERR_NULL (-1)
#define CG_ERR_SHAPE (-2)
#define CG_ERR_DTYPE (-3)
#define CG_ERR_ALIGN (-4)
#define CG_ERR_WORKSPACE (-5)
#define CG_ERR_MODE (-6)
#define CG_ERR_UNSUPPORTED (-7)

typedef struct CUstream_st* cg_stream_t; /* == cudaStream_t */

int cg_abi_version(void);
const char* cg_status_string(int status);

/* Bytes of scratch cg_rglru_fwd / cg_rnn_scan_fwd need for a [B,T,E] problem
 * (upper bound over all arithmetic modes and kernel geometries).  The scratch
 * must be ZERO-FILLED ONCE before its first use (cudaMemset / torch.zeros); it
 * is self-cleaning afterwards and may be reused by any later call, of any
 * shape that fits, issued on the same stream. */
size_t cg_scan_workspace_bytes(int B, int T, int E, int dtype);

/*
 * Conv1D prefill (cache is None): layers.py:484-546.
 *   x [B,T,E], w [W,E], b [E], y [B,T,E] in `dtype`;
 *   seg: segment positions, int32 or int64 (seg_is_i64), row b at
 *        seg + b*seg_batch_stride elements (0 broadcasts one row, :512-513);
 *   cache_out (nullable): [B,W-1,E] in `dtype` -- the last W-1 input rows, left
 *        zero padded (:542-543).
 * W == 4 runs the vectorised kernel; any other W >= 1 a generic one.
 */
int cg_conv1d_fwd(const void* x, const void* w, const void* b, const void* seg,
                  int seg_is_i64, long long seg_batch_stride, void* y,
                  void* cache_out, int B, int T, int E, int W, int dtype,
                  int mask_mode, int arith_mode, cg_stream_t stream);

/*
 * Conv1D decode step (cache given, T == 1, no mask): layers.py:478-483, :542.
 *   x [B,1,E], cache_in / cache_out [B,W-1,E] in cache_dtype (fp32 or bf16;
 *   cache_out may alias cache_in), y [B,1,E].
 */
int cg_conv1d_decode(const void* x, const void* w, const void* b,
                     const void* cache_in, int cache_dtype, void* y,
                     void* cache_out, int B, int E, int W, int dtype,
                     int arith_mode, cg_stream_t stream);

/*
 * RG-LRU after the gate GEMMs: layers.py:345-375 (gate math + rnn_scan).
 *   x       [B,T,E]  Conv1D output
 *   gemm_x  [B,T,E]  input_gate GEMM output WITHOUT bias  } rows gate_row_stride
 *   gemm_a  [B,T,E]  a_gate GEMM output WITHOUT bias      } elements apart
 *   gate_block_width  0: both are plain rows, channel c at column c.
 *        bw > 0 (= E / num_heads): ONE fused block-diagonal GEMM wrote rows of
 *        [num_heads][2*bw] (per head: bw input-gate then bw a-gate columns);
 *        channel c then sits at column c + (c / bw) * bw of gemm_x, and gemm_a
 *        must point bw elements past gemm_x (gate_row_stride >= 2E).
 *   bias_x, bias_a  [E] or NULL: pre = round(gemm + bias) (layers.py:139)
 *   a_param [E]
 *   h0      [B,E] fp32 or NULL;  last_h [B,E] fp32 or NULL
 *   y       [B,T,E]
 * T == 1 follows rnn_scan's sampling branch (:175-182).
 */
int cg_rglru_fwd(const void* x, const void* gemm_x, const void* gemm_a,
                 long long gate_row_stride, int gate_block_width,
                 const void* bias_x, const void* bias_a, const void* a_param,
                 const void* seg,
                 int seg_is_i64, long long seg_batch_stride, const float* h0,
                 void* y, float* last_h, void* workspace,
                 size_t workspace_bytes, int B, int T, int E, int dtype,
                 int arith_mode, cg_stream_t stream);

/*
 * rnn_scan: layers.py:146-199.  x, a, y [B,T,E] in `dtype`; reset [B,T] uint8
 * (non-zero = reset); h0 / last_h as above.  arith_mode: 0 = chunked two-level
 * scan, CG_ARITH_STRICT = sequential non-FMA kernel (bit-exact with the
 * reference loop).
 */
int cg_rnn_scan_fwd(const void* x, const void* a, const unsigned char* reset,
                    const float* h0, void* y, float* last_h, void* workspace,
                    size_t workspace_bytes, int B, int T, int E, int dtype,
                    int arith_mode, cg_stream_t stream);

/*
 * Backward of the Conv1D prefill (W == 4; other widths: CG_ERR_UNSUPPORTED): the
 * gradients autograd derives from layers.py:484-546 with the forward's
 * document mask m_s(t) of tap s at output step t (mask_mode as cg_conv1d_fwd):
 *     dx[u]   = sum_s w[3-s] * gy[u+s] * m_s(u+s)
 *     dw[3-s] = sum_{b,t} gy[t] * x[t-s] * m_s(t)          db = sum_{b,t} gy[t]
 * dx is accumulated in `dtype` in autograd's order (tap 3 first, every product
 * and sum rounded: bit-exact with the reference); dw / db are summed in fp32 in
 * a fixed order (bit-reproducible) and rounded once.  gy, x, dx [B,T,E]; w, dw [4,E]; db [E].
 * workspace: cg_conv1d_bwd_workspace_bytes(B,T,E) bytes, no initialisation needed.
 * SURVEY.md section 8(f) row F4 (training path).
 */
size_t cg_conv1d_bwd_workspace_bytes(int B, int T, int E);
int cg_conv1d_bwd(const void* gy, const void* x, const void* w, const void* seg,
                  int seg_is_i64, long long seg_batch_stride, void* dx, void* dw,
                  void* db, void* workspace, size_t workspace_bytes, int B, int T,
                  int E, int W, int dtype, int mask_mode, cg_stream_t stream);

/*
 * Training path of the RG-LRU gate math.
 *   cg_rglru_gates_fwd: layers.py:348-365 from x and the gate pre-activations
 *     (BlockDiagonalLinear outputs, bias included) to the scan's inputs
 *     a = exp(-8 sigmoid(pre_a) softplus(a_param)) and
 *     nx = x sigmoid(pre_x) sqrt(1 - a^2) (multiplier 1 where reset != 0), with the
 *     reference's eager bf16 rounding points.  All [B,T,E] in `dtype`, reset [B,T].
 *   cg_rglru_gates_bwd: what autograd derives from those lines, including the
 *     clipped square-root gradient (layers.py:224-238): from d_nx = grad(nx) and
 *     d_a = grad(a) (the outputs of cg_rnn_scan_bwd) to dx, d_pre_x, d_pre_a
 *     [B,T,E] and d_a_param [E] (fixed-order reduction).  fp32 in registers, gates
 *     recomputed from the saved pre-activations.
 */
int cg_rglru_gates_fwd(const void* x, const void* pre_x, const void* pre_a, const void* a_param,
                       const unsigned char* reset, void* a, void* nx, int B, int T, int E,
                       int dtype, cg_stream_t stream);
size_t cg_rglru_gates_bwd_workspace_bytes(int B, int T, int E);
int cg_rglru_gates_bwd(const void* x, const void* pre_x, const void* pre_a, const void* a_param,
                       const unsigned char* reset, const void* d_nx, const void* d_a, void* dx,
                       void* d_pre_x, void* d_pre_a, void* d_a_param, void* workspace,
                       size_t workspace_bytes, int B, int T, int E, int dtype, cg_stream_t stream);

/*
 * Backward of rnn_scan: what torch autograd derives from the reference loop
 * (layers.py:173, :187-199) -- SURVEY.md section 8(f) row F4; the JAX VJP
 * jax/pallas.py:785-839 is the same recurrence.  One chunked scan over
 * reversed time:
 *     dh_t = a_{t+1} * ~reset_{t+1} * dh_{t+1} + gy_t      (dh_T = g_last_h)
 *     dx_t = dh_t        da_t = ~reset_t * dh_t * h_{t-1}  (h_{-1} = h0)
 *     dh0  = a_0 * ~reset_0 * dh_0
 * fp32 mul then add as autograd does; dx / da leave in `dtype`.
 *   gy [B,T,E] grad of y (`dtype`); g_last_h [B,E] fp32 grad of last_h or NULL;
 *   a, reset, h0 as given to the forward; h [B,T,E] = the forward's y (`dtype`;
 *   for bf16 the reference keeps fp32 h_t for this product, so da differs from
 *   it by at most the bf16 rounding of h);
 *   dx, da [B,T,E] in `dtype`; dh0 [B,E] fp32 or NULL.
 * Workspace as cg_rnn_scan_fwd.
 */
int cg_rnn_scan_bwd(const void* gy, const float* g_last_h, const void* a, const void* h,
                    const unsigned char* reset, const float* h0, void* dx, void* da,
                    float* dh0, void* workspace, size_t workspace_bytes, int B, int T,
                    int E, int dtype, cg_stream_t stream);

/*
 * Fused tensor-core RG-LRU (bf16): the two BlockDiagonalLinear gate GEMMs
 * (layers.py:133-142, :348-349), the gate math (:350-365) and rnn_scan
 * (:146-199, :366-371) in ONE kernel -- tcgen05 MMAs with TMEM accumulators,
 * TMA operand loads, gate math + scan in the epilogue; the pre-activations
 * never reach HBM.  SURVEY.md section 8(f) row F1.
 *
 *   cg_rglru_fused_supported   1 if (E, H, dtype) can take this path: bf16 and a
 *                              head width E/H of 128 or 256; else use cg_rglru_fwd.
 *   cg_rglru_gate_pack_bytes   size of the packed gate-weight buffer.
 *   cg_rglru_pack_gate_weights wx, wa [H, bw, bw] (reference layout, layers.py:103:
 *                              y = x @ w[h]) -> wpack: the shared-memory image
 *                              (K-major, 128 B swizzle) the MMAs read.  Call once
 *                              per weight update.
 *   cg_rglru_fused_workspace_bytes  scratch size; zero-fill ONCE before first use
 *                              (same rules as cg_scan_workspace_bytes).  Every wait inside
 *                              the kernel is bounded by a clock watchdog; if one ever
 *                              expires (a protocol error) the kernel TRAPS: the stream
 *                              carries a sticky CUDA error and the next synchronising
 *                              call of the host fails -- results are never silently wrong.
 *   cg_rglru_fused_fwd         x [B,T,E] Conv1D output; biases [E] or NULL;
 *                              h0 / last_h / y / seg as in cg_rglru_fwd.
 *                              arith_mode: CG_ARITH_REFERENCE or CG_ARITH_FAST
 *                              (every eager bf16 rounding point reproduced in both;
 *                              the GEMM accumulates in fp32 and rounds once to bf16
 *                              like the reference's einsum); CG_ARITH_VARIANT(v), v in 1..64 limits
 *                              the launch to v CTAs (a test hook: fewer CTAs than column
 *                              families exercises the per-CTA family loop).
 *                              gate_mul (nullable): [B,T,E] bf16; when given, the kernel
 *                              returns round_bf16(y * gate_mul) -- the gating product of
 *                              RecurrentBlock.forward (modules.py:651, `x = x * y`) folded
 *                              into the store; last_h is
 *                              unaffected.
 */
int cg_rglru_fused_supported(int E, int H, int dtype);
size_t cg_rglru_gate_pack_bytes(int E, int H);
int cg_rglru_pack_gate_weights(const void* wx, const void* wa, void* wpack,
                               int E, int H, int dtype, cg_stream_t stream);
size_t cg_rglru_fused_workspace_bytes(int B, int T, int E);
int cg_rglru_fused_fwd(const void* x, const void* wpack, const void* bias_x,
                       const void* bias_a, const void* a_param, const void* seg,
                       int seg_is_i64, long long seg_batch_stride,
                       const float* h0, void* y, float* last_h, void* workspace,
                       size_t workspace_bytes, int B, int T, int E, int H,
                       int dtype, int arith_mode, const void* gate_mul,
                       cg_stream_t stream);

/*
 * Conv1D.forward -> RGLRU.forward of one recurrent block as ONE kernel (prefill,
 * bf16, W == 4, head width 128 / 256): what RecurrentBlock.forward does between
 * its projections (recurrentgemma/torch/modules.py:638-649) --
 *     x, conv1d_state = conv_1d(x, segment_pos)       layers.py:458-546
 *     x, rg_lru_state = rg_lru(x, segment_pos)        layers.py:322-375
 * -- with the temporal convolution computed INSIDE the fused tensor-core kernel:
 * TMA loads the rows of x (= linear_x output) into shared memory, the epilogue
 * warpgroups convolve them in place with the reference's rounding (bit-exact with
 * cg_conv1d_fwd; at head width 256 once per head: the two CTAs of a head form a
 * cluster and exchange their halves), the tcgen05 gate GEMMs read the result,
 * gate math and scan follow in the epilogue.  Neither the conv output nor the gate
 * pre-activations reach HBM: the step moves 2 x B*T*E*2 bytes (x in, y out).
 * Faster than cg_conv1d_fwd + cg_rglru_fused_fwd at every measured shape; same bits.
 * H <= 32 (CG_ERR_UNSUPPORTED otherwise).
 *   x [B,T,E] Conv1D INPUT; conv_w [4,E]; conv_b [E];
 *   conv_cache_out (nullable) [B,3,E] bf16: the last three rows of x, left zero
 *        padded (layers.py:542-543);
 *   everything else as cg_rglru_fused_fwd (wpack from cg_rglru_pack_gate_weights,
 *   workspace of cg_rglru_fused_workspace_bytes, gate_mul nullable).
 *   mask_mode: CG_MASK_FORK / CG_MASK_UPSTREAM.  W must be 4 (CG_ERR_UNSUPPORTED
 *   otherwise: run cg_conv1d_fwd + cg_rglru_fused_fwd).
 */
int cg_recurrent_prefill_fwd(const void* x, const void* conv_w, const void* conv_b,
                             const void* wpack, const void* bias_x, const void* bias_a,
                             const void* a_param, const void* seg, int seg_is_i64,
                             long long seg_batch_stride, const float* h0,
                             const void* gate_mul, void* y, void* conv_cache_out,
                             float* last_h, void* workspace, size_t workspace_bytes,
                             int B, int T, int E, int H, int W, int dtype,
                             int mask_mode, int arith_mode, cg_stream_t stream);

/*
 * Fused decode step of the recurrent hot path (T == 1, caches given, bf16,
 * W == 4, head width a multiple of 64 up to 256): the Conv1D step with its cache
 * roll (layers.py:478-483, :542), both BlockDiagonalLinear gate GEMVs (:133-142),
 * the gate math (:345-365) and rnn_scan's sampling branch h = a*h0 + x~
 * (:175-182) in ONE launch instead of three (conv step, cuBLAS batched GEMV,
 * gate step).  SURVEY.md section 8(b) / 8(f) row F3.
 *   x [B,1,E] = linear_x output; conv_w [4,E], conv_b [E];
 *   cache_in / cache_out [B,3,E] in cache_dtype (cache_out nullable, must NOT
 *   alias cache_in); wx, wa [H,bw,bw] in the reference layout (y = x @ w[h]);
 *   h0 [B,E] fp32 or NULL; gate_mul [B,1,E] or NULL (modules.py:651 folded in);
 *   y [B,1,E]; last_h [B,E] fp32 or NULL.
 *   arith_mode: CG_ARITH_REFERENCE or CG_ARITH_FAST.
 */
int cg_recurrent_decode_supported(int E, int H, int W, int dtype);
int cg_recurrent_decode_step(const void* x, const void* conv_w, const void* conv_b,
                             const void* cache_in, int cache_dtype,
                             const void* wx, const void* wa, const void* bias_x,
                             const void* bias_a, const void* a_param,
                             const void* seg, int seg_is_i64,
                             long long seg_batch_stride, const float* h0,
                             const void* gate_mul, void* y, void* cache_out,
                             float* last_h, int B, int E, int H, int W, int dtype,
                             int arith_mode, cg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CADENCE_B200_H_ */
