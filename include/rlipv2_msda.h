/*
 * rlipv2_msda.h - C ABI of the B200-native multi-scale deformable attention op.
 *
 * This is the drop-in boundary for the one native component of RLIPv2: the python
 * extension module `MultiScaleDeformableAttention` that the reference binds with pybind11 in
 *   /root/reference/models/ops/src/vision.cpp:13-16           (module definition)
 *   /root/reference/models/ops/src/ms_deform_attn.h:20-61     (device dispatch)
 *   /root/reference/models/ops/src/cuda/ms_deform_attn_cuda.cu:20-80   (forward host wrapper)
 *   /root/reference/models/ops/src/cuda/ms_deform_attn_cuda.cu:83-153  (backward host wrapper)
 * and calls from models/ops/functions/ms_deform_attn_func.py:27-44.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to a dense, C-contiguous array (the reference asserts
 *     contiguity, ms_deform_attn_cuda.cu:28-38); `stream` is a cudaStream_t passed as void*
 *     (NULL = legacy default stream).  Calls are asynchronous: no host synchronisation.
 *   - tensor layouts are the reference's:
 *        value            [batch, spatial_size, num_heads, channels]
 *        spatial_shapes   [num_levels, 2]  int64 (H_l, W_l), read on the device
 *        level_start_idx  [num_levels]     int64, read on the device
 *        sampling_loc     [batch, num_query, num_heads, num_levels, num_point, 2]  (x=w, y=h) in [0,1]
 *        attn_weight      [batch, num_query, num_heads, num_levels, num_point]
 *        out / grad_out   [batch, num_query, num_heads*channels]
 *   - return value: 0 on success; a positive value is a cudaError_t from the launch (the
 *     reference only printf()s launch errors, ms_deform_im2col_cuda.cuh:948-952 - here they are
 *     returned); a negative value is one of RLIPV2_MSDA_E*.
 *   - re-entrant, no global mutable state; the im2col_step batching loop of the reference
 *     (ms_deform_attn_cuda.cu:50-75) is not needed: one launch covers the whole batch, so the
 *     python binding only validates `batch % min(batch, im2col_step) == 0` for error parity.
 *
 * There is deliberately NO CPU implementation behind this ABI (the reference has none either:
 * cpu/ms_deform_attn_cpu.cpp:17-41 raises).
 */
#ifndef RLIPV2_MSDA_H_
#define RLIPV2_MSDA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLIPV2_MSDA_ABI_VERSION 2

#define RLIPV2_MSDA_EINVAL   (-1) /* negative / zero dimension where not allowed, null pointer */
#define RLIPV2_MSDA_ETOOBIG  (-2) /* an index would overflow the 64-bit-safe limits we support  */
#define RLIPV2_MSDA_ESHAPE   (-3) /* fused-prologue entry points: shape outside fp32, D=32, L=4, P=4 */
#define RLIPV2_MSDA_EALIGN   (-4) /* fused-prologue entry points: a pointer is not 16-byte aligned (the plain entry
                                     points fall back to the generic kernel instead, like the reference's scalar one) */

/* ms_deform_attn_cuda_forward (ms_deform_attn_cuda.cu:20-80), scalar_t = float.
 * Writes every element of `out` (no pre-zeroing needed). */
int rlipv2_msda_forward_f32(const float *value, const int64_t *spatial_shapes,
                            const int64_t *level_start_index, const float *sampling_loc,
                            const float *attn_weight, int batch, int spatial_size, int num_heads,
                            int channels, int num_levels, int num_query, int num_point,
                            float *out, void *stream);

/* same, scalar_t = double (AT_DISPATCH_FLOATING_TYPES, ms_deform_attn_cuda.cu:64) */
int rlipv2_msda_forward_f64(const double *value, const int64_t *spatial_shapes,
                            const int64_t *level_start_index, const double *sampling_loc,
                            const double *attn_weight, int batch, int spatial_size, int num_heads,
                            int channels, int num_levels, int num_query, int num_point,
                            double *out, void *stream);

/* ms_deform_attn_cuda_backward (ms_deform_attn_cuda.cu:83-153), scalar_t = float.
 * All three gradient arrays are fully overwritten: grad_value is zero-filled on `stream` and then
 * accumulated into (fp32 reductions, order not deterministic - same as the reference's atomicAdd,
 * ms_deform_im2col_cuda.cuh:125-152); grad_sampling_loc and grad_attn_weight are written once. */
int rlipv2_msda_backward_f32(const float *value, const int64_t *spatial_shapes,
                             const int64_t *level_start_index, const float *sampling_loc,
                             const float *attn_weight, const float *grad_out, int batch,
                             int spatial_size, int num_heads, int channels, int num_levels,
                             int num_query, int num_point, float *grad_value,
                             float *grad_sampling_loc, float *grad_attn_weight, void *stream);

int rlipv2_msda_backward_f64(const double *value, const int64_t *spatial_shapes,
                             const int64_t *level_start_index, const double *sampling_loc,
                             const double *attn_weight, const double *grad_out, int batch,
                             int spatial_size, int num_heads, int channels, int num_levels,
                             int num_query, int num_point, double *grad_value,
                             double *grad_sampling_loc, double *grad_attn_weight, void *stream);

/* ---- fused-prologue variant (the module's arithmetic folded into the op) -----------------------------
 * Replaces, for 2-d reference points (the encoder self-attention; `reference_points.shape[-1] == 2`),
 *   /root/reference/models/ops/modules/ms_deform_attn.py:102-109
 *       attention_weights = F.softmax(attention_weights(query).view(N, Lq, M, L*P), -1)
 *       sampling_locations = reference_points[:, :, None, :, None, :]
 *                            + sampling_offsets / offset_normalizer[None, None, None, :, None, :]
 *   + the op call of :116 - i.e. the ~6 elementwise kernels over [N, Lq, M, L, P, 2] tensors that the
 *   reference (through torch) launches around the native op, forward and backward.
 *     reference_points [batch, num_query, num_levels, 2]   (x, y) in [0,1], not differentiated
 *     proj             [batch, num_query, num_heads*num_levels*num_point*3]: the query projected by the
 *                      stacked weights [sampling_offsets.weight ; attention_weights.weight]: first the
 *                      M*L*P*2 raw sampling offsets, then the M*L*P raw attention logits
 *     grad_proj        same shape; every element written once (no pre-zeroing needed)
 * Only fp32, channels = 32, num_levels = 4, num_point = 4 (every ParSeDA call); other shapes return
 * RLIPV2_MSDA_ESHAPE and the caller uses the plain entry points. */
int rlipv2_msda_proj_forward_f32(const float *value, const int64_t *spatial_shapes,
                                 const int64_t *level_start_index, const float *reference_points,
                                 const float *proj, int batch, int spatial_size, int num_heads, int channels,
                                 int num_levels, int num_query, int num_point, float *out, void *stream);

int rlipv2_msda_proj_backward_f32(const float *value, const int64_t *spatial_shapes,
                                  const int64_t *level_start_index, const float *reference_points,
                                  const float *proj, const float *grad_out, int batch, int spatial_size,
                                  int num_heads, int channels, int num_levels, int num_query, int num_point,
                                  float *grad_value, float *grad_proj, void *stream);

/* The same fused prologue for 4-d reference points (the decoders' cross-attention, anchors (cx, cy, w, h) per level,
 * ms_deform_attn.py:110-112): reference_boxes [batch, num_query, num_levels, 4], not differentiated;
 *     loc = box[:2] + offset / num_point * box[2:] * 0.5
 * and in the backward d offset = d loc * 0.5 * box[2:] / num_point. */
int rlipv2_msda_proj_ref4_forward_f32(const float *value, const int64_t *spatial_shapes,
                                      const int64_t *level_start_index, const float *reference_boxes,
                                      const float *proj, int batch, int spatial_size, int num_heads, int channels,
                                      int num_levels, int num_query, int num_point, float *out, void *stream);

int rlipv2_msda_proj_ref4_backward_f32(const float *value, const int64_t *spatial_shapes,
                                       const int64_t *level_start_index, const float *reference_boxes,
                                       const float *proj, const float *grad_out, int batch, int spatial_size,
                                       int num_heads, int channels, int num_levels, int num_query, int num_point,
                                       float *grad_value, float *grad_proj, void *stream);

/* Human-readable text for a return code of the functions above (static storage). */
/* Experimental forward variant with north_star's literal design: the cells of the two coarsest levels of one (image, head) -
 * the cell range [coarse_start, spatial_size), coarse_start = level_start_index[num_levels - 2] passed as a HOST integer -
 * are staged in shared memory by TMA (cp.async.bulk.tensor 3-D boxes) and sampled from there; levels 0 / 1 are gathered from
 * global memory.  Same arguments / result as rlipv2_msda_forward_f32 (fp32, D = 32, L = 4, P = 4 only;
 * RLIPV2_MSDA_ESHAPE when the coarse levels exceed one SM's shared memory).  Measured slower than the default kernel
 * (profiles/msda_r02.md); kept for the comparison. */
int rlipv2_msda_forward_tma_f32(const float *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                                const float *sampling_loc, const float *attn_weight, int batch, int spatial_size,
                                int num_heads, int channels, int num_levels, int num_query, int num_point,
                                int coarse_start, float *out, void *stream);

const char *rlipv2_msda_error_string(int code);

/* RLIPV2_MSDA_ABI_VERSION the library was built with. */
int rlipv2_msda_abi_version(void);

/* Backward schedule of the fast path (fp32, D = 32, L = 4, P = 4), process-wide:
 *   0  one red.global.add.v4.f32 per valid corner - the reference's one atomicAdd per corner and channel
 *      (ms_deform_im2col_cuda.cuh:125,134,143,152), vectorised;
 *   1  corners of one (image, query, head) that fall on the same cell of a level are merged before they are issued
 *      (every contribution is a scalar times the pair's grad_out row, so the scalars add), and corners whose merged scalar
 *      is exactly zero are not issued.  Same sums as mode 0 up to fp32 rounding order (rlipv2_b200/csrc/msda_merge.h);
 *   2  (default) by call shape: mode 1 for calls with at least 8192 (image, query) rows - the encoder's self-attention, whose
 *      points cluster around the query's own cell; for the rest (decoder-shaped calls: nothing merges, DRAM-latency-bound) the
 *      unmerged schedule, which rlipv2_msda_backward_f32 then runs at 3 CTAs per SM instead of 2 (same source, 80 registers);
 *   3 / 4  measurement only: the merged / unmerged schedule of rlipv2_msda_backward_f32 forced to 3 CTAs per SM.
 * Returns 0, or RLIPV2_MSDA_EINVAL for another mode. */
int rlipv2_msda_set_backward_mode(int mode);
int rlipv2_msda_get_backward_mode(void);

/* Number of kernels launched by this library in this process since load (for bench.py's
 * `gpu_launches`; relaxed atomic counter, never reset). */
unsigned long long rlipv2_msda_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* RLIPV2_MSDA_H_ */
