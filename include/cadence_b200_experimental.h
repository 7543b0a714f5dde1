/*
 * cadence_b200_experimental.h -- development taps of libcadence_b200.so.
 *
 * NOT part of the drop-in surface (include/cadence_b200.h is): these entry points
 * exist for the repo's own tests and profiling scripts, may change without an ABI
 * version bump and need not be bound by a maintainer of the reference.
 */
#ifndef CADENCE_B200_EXPERIMENTAL_H_
#define CADENCE_B200_EXPERIMENTAL_H_

#include "cadence_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/*
 * cg_recurrent_prefill_fwd / cg_rglru_fused_fwd (conv_w == NULL: x is the Conv1D
 * output) with a debug tap: debug_out [3][B][T][E] bf16 receives, per element, the
 * rounded input-gate pre-activation, the rounded a-gate pre-activation and the
 * (convolved) x exactly as the epilogue of the fused kernel saw them
 * (tests compare them with cuBLAS / cg_conv1d_fwd).  No gate_mul.
 */
int cg_recurrent_prefill_debug(const void* x, const void* conv_w, const void* conv_b,
                               const void* wpack, const void* bias_x, const void* bias_a,
                               const void* a_param, const void* seg, int seg_is_i64,
                               long long seg_batch_stride, const float* h0, void* y,
                               void* conv_cache_out, float* last_h, void* workspace,
                               size_t workspace_bytes, int B, int T, int E, int H,
                               int dtype, int mask_mode, int arith_mode, void* debug_out,
                               cg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CADENCE_B200_EXPERIMENTAL_H_ */
