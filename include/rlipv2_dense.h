/*
 * rlipv2_dense.h - C ABI of the tcgen05 (sm_100a tensor-core) dense contraction used by the
 * ParSeDA hot path: y[M,N] = act(x[M,K] . W[N,K]^T + bias[N]).
 *
 * It replaces, for the `nn.Linear` calls of
 *   /root/reference/models/fuse_helper.py:370-373,463-464            (ALIF projections)
 *   /root/reference/models/modeling_roberta.py:159-165,252,318,332   (RobertaLayer)
 *   /root/reference/models/dab_deformable/deformable_transformer.py:1283-1287,1368-1372 (FFNs)
 *   /root/reference/models/ops/modules/ms_deform_attn.py:98,102,103,118 (MSDeformAttn projections)
 * the cuBLAS SGEMM + bias + activation kernels torch launches (F.linear / F.relu / F.gelu).
 *
 * Conventions: device pointers to dense row-major fp32 arrays, 16-byte aligned; `stream` is a
 * cudaStream_t passed as void*; asynchronous; returns 0, a positive cudaError_t, or RLIPV2_DENSE_E*.
 * Products are TF32 (10-bit mantissa), accumulation fp32.  No CPU implementation exists.
 */
#ifndef RLIPV2_DENSE_H_
#define RLIPV2_DENSE_H_

#ifdef __cplusplus
extern "C" {
#endif

#define RLIPV2_DENSE_ACT_NONE 0
#define RLIPV2_DENSE_ACT_RELU 1
#define RLIPV2_DENSE_ACT_GELU 2   /* exact erf GELU, as F.gelu / HF ACT2FN["gelu"] */

#define RLIPV2_DENSE_EINVAL  (-1)
#define RLIPV2_DENSE_ESHAPE  (-2)  /* needs N % 128 == 0 and K % 32 == 0 */
#define RLIPV2_DENSE_EALIGN  (-3)
#define RLIPV2_DENSE_EDRIVER (-4)

/* 1 if (M, N, K) can run on the tcgen05 kernel, else 0 (callers route other shapes to cuBLAS) */
int rlipv2_dense_linear_tf32_supported(int M, int N, int K);

/* y = act(x . w^T + bias); x [M,K], w [N,K] (nn.Linear layout), bias [N] or NULL, y [M,N] */
int rlipv2_dense_linear_tf32(const float *x, const float *w, const float *bias, float *y, int M, int N, int K,
                             int act, void *stream);

/* Backward contractions of the same nn.Linear (what autograd's mm / addmm backward runs through cuBLAS):
 *
 * weight gradient  dw[N,K] += g[T,N]^T . x[T,K]   (g = dL/dy, x = the layer input; T = rows = tokens).
 *   Both operands are read as stored (MN-major tcgen05 operands, no transposed copies); the T axis is split over
 *   `splits` CTAs per output tile and the partial tiles are reduced into `dw` with red.global.add.v4.f32, so
 *   `dw` must be initialised by the caller (zeros, or the gradient accumulated so far) and the fp32 summation
 *   order is not deterministic.  N % 4 == 0, K % 4 == 0.
 *
 * input gradient   dx[T,K] = g[T,N] . w[N,K]      (w = nn.Linear weight as stored).
 *   With relu_out != NULL (the [T,K] output of the ReLU layer that feeds this Linear) the ReLU backward and the
 *   bias gradient of that previous layer are fused into the epilogue: dx = (relu_out > 0) ? dx : 0 and
 *   colsum[ceil(T/128), K] = column sums of the masked dx per 128-row tile (fully written, no atomics; the caller
 *   adds the tiles up).  N % 4 == 0, K % 4 == 0. */
int rlipv2_dense_wgrad_tf32(const float *g, const float *x, float *dw, int T, int N, int K, int splits, void *stream);
int rlipv2_dense_dgrad_tf32(const float *g, const float *w, float *dx, const float *relu_out, float *colsum,
                            int T, int N, int K, void *stream);

/* rlipv2_dense_linear_tf32 with `value.masked_fill(padding_mask[..., None], 0)`
 * (/root/reference/models/ops/modules/ms_deform_attn.py:98-100) folded into the epilogue: rows r with
 * rowmask[r] != 0 are written as zeros.  rowmask: M bytes (a torch bool tensor) or NULL. */
int rlipv2_dense_linear_tf32_rowmask(const float *x, const float *w, const float *bias, const unsigned char *rowmask,
                                     float *y, int M, int N, int K, int act, void *stream);

/* Split-K variant of rlipv2_dense_linear_tf32 (no activation) for the small-M / long-K linears of ALIF and the RobertaLayer
 * (/root/reference/models/fuse_helper.py:463-464 out_v_proj / out_l_proj with K = 2048, :370-373 the label-side
 * in-projections with K = 768, models/modeling_roberta.py:332 FFN-down with K = 3072): with a few hundred rows these fill
 * 4-64 of 148 SMs with one CTA per output tile, so the K axis is split over `splits` CTAs per tile whose partial tiles are
 * reduced into y with fp32 reductions (y is zero-filled by the call; the first slice adds the bias).  The summation order
 * over K slices is not fixed between runs (fp32 rounding level).  K % 32 == 0, N % 4 == 0. */
int rlipv2_dense_linear_splitk_tf32(const float *x, const float *w, const float *bias, float *y, int M, int N, int K,
                                    int splits, void *stream);

/* Tuning knob of rlipv2_dense_linear_tf32* for grids that leave SMs idle (M of a few hundred rows: the ALIF,
 * RobertaLayer and decoder linears).  0: always 128x128 tiles with a 3-stage TMA ring (2 CTAs/SM); 1: grids of
 * at most one CTA per SM use a 6-stage ring; 2 (default): additionally 128x64 tiles with an 8-stage ring while the grid
 * still fits one CTA per SM.  Results are identical in every mode (same products, same fp32 accumulation order per output). */
void rlipv2_dense_set_small_mode(int mode);
int rlipv2_dense_get_small_mode(void);

/* Large grids: linears with more than `tiles` 128 x 128 output tiles run the PERSISTENT kernel (one CTA per SM walks the
 * tiles, TMA ring across tile boundaries, accumulator double-buffered in tensor memory so that a tile's epilogue overlaps the
 * next tile's main loop); 0 = never.  The encoder's 44 446-row linears have thousands of tiles of only 8 k-blocks each.
 * Results are identical to the one-tile-per-CTA kernel (same products, same accumulation order per output). */
void rlipv2_dense_set_persistent_min_tiles(int tiles);
int rlipv2_dense_get_persistent_min_tiles(void);

/* The same schedule for rlipv2_dense_dgrad_tf32 with a ReLU gate (dgrad_mask_persistent_kernel, 128 x 256 tiles); applies to
 * problems above the tile threshold when switched on.  Off by default: in the train step the gated input gradient runs
 * beside the weight-gradient kernels of the parameter-gradient stream, and a one-CTA-per-SM kernel that takes all shared
 * memory keeps them from co-running (measured r02l: 26.9 vs 26.5 ms/step). */
void rlipv2_dense_set_persistent_dgrad(int on);
int rlipv2_dense_get_persistent_dgrad(void);

const char *rlipv2_dense_error_string(int code);

/* kernels launched by this library in this process (for bench.py's gpu_launches) */
unsigned long long rlipv2_dense_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* RLIPV2_DENSE_H_ */
