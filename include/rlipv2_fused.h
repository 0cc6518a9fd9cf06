/*
 * rlipv2_fused.h - C ABI of the HBM-bound fused kernels around the dense contractions of the ParSeDA
 * train step ("fused deformable-layer epilogues").  Each replaces a chain
 * of torch kernels that the reference's modules launch separately:
 *
 *   add + LayerNorm forward / backward   src = norm(src + dropout(src2)) in
 *       /root/reference/models/dab_deformable/deformable_transformer.py:1295-1296, 1285-1286, 1390-1400,
 *       models/modeling_roberta.py:252-256, 332-336 (F.layer_norm + aten::add; backward:
 *       layer_norm_grad_input + GammaBetaBackward)
 *   ReLU-mask + bias-gradient column sum  the backward of `self.activation(self.linear1(src))`
 *       (deformable_transformer.py:1284) and of every nn.Linear bias (aten::threshold_backward,
 *       aten::sum over rows)
 *   flat AdamW step                       torch.optim.AdamW of main.py:538-539 over one flat
 *       parameter / gradient / moment buffer per learning-rate group
 *
 * Conventions as in rlipv2_msda.h: dense row-major fp32 device arrays, `stream` = cudaStream_t as
 * void*, asynchronous, return 0 / positive cudaError_t / negative RLIPV2_FUSED_E*.  No CPU versions.
 */
#ifndef RLIPV2_FUSED_H_
#define RLIPV2_FUSED_H_

#ifdef __cplusplus
extern "C" {
#endif

#define RLIPV2_FUSED_EINVAL (-1)
#define RLIPV2_FUSED_ESHAPE (-2)   /* LayerNorm kernels: C must be a multiple of 128 and <= 1024 */

/* y = LayerNorm(x + r) * gamma + beta over the last dim (C); also writes z = x + r, mean[M], rstd[M]
 * for the backward.  r may be NULL (plain LayerNorm of x; z is still written). */
int rlipv2_add_layernorm_fwd_f32(const float *x, const float *r, const float *gamma, const float *beta,
                                 float eps, int M, int C, float *y, float *z, float *mean, float *rstd,
                                 void *stream);

/* dz = d(LayerNorm)/dz (the gradient of both x and r), dgamma[C] and dbeta[C] (overwritten). */
int rlipv2_layernorm_bwd_f32(const float *dy, const float *z, const float *mean, const float *rstd,
                             const float *gamma, int M, int C, float *dz, float *dgamma, float *dbeta,
                             void *stream);

/* colsum[N] = sum over rows of g [M,N] (overwritten).  If y != NULL the ReLU mask is applied first:
 * gm = (y > 0) ? g : 0, gm is written to `gmasked` (may alias g) and summed instead. */
int rlipv2_relu_bwd_colsum_f32(const float *g, const float *y, float *gmasked, float *colsum, int M, int N,
                               void *stream);

/* Backward of `y.masked_fill(rowmask[..., None], 0)` (ms_deform_attn.py:99-100) fused with the bias gradient:
 * gmasked[r,:] = rowmask[r] ? 0 : g[r,:] (may alias g), colsum[N] = column sums of gmasked (overwritten).
 * rowmask: M bytes (a torch bool tensor). */
int rlipv2_rowmask_bwd_colsum_f32(const float *g, const unsigned char *rowmask, float *gmasked, float *colsum, int M,
                                  int N, void *stream);

/* `_acc` variants: with accumulate != 0 the reduced outputs (dgamma/dbeta, colsum) are ADDED to what the arrays hold
 * instead of overwriting them - the parameter-gradient accumulation (`param.grad += ...`, autograd's AccumulateGrad
 * in the reference's loss.backward(), engine.py:163) folded into the producing kernel, writing straight into the
 * flat gradient buffer. */
int rlipv2_layernorm_bwd_acc_f32(const float *dy, const float *z, const float *mean, const float *rstd,
                                 const float *gamma, int M, int C, float *dz, float *dgamma, float *dbeta,
                                 int accumulate, void *stream);
int rlipv2_relu_bwd_colsum_acc_f32(const float *g, const float *y, float *gmasked, float *colsum, int M, int N,
                                   int accumulate, void *stream);
int rlipv2_rowmask_bwd_colsum_acc_f32(const float *g, const unsigned char *rowmask, float *gmasked, float *colsum,
                                      int M, int N, int accumulate, void *stream);

/* One AdamW step over n contiguous elements (decoupled weight decay, torch.optim.AdamW semantics).
 * `step` is a device float holding the 1-based step count of THIS update (bias correction). */
int rlipv2_adamw_f32(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, long long n,
                     double lr, double beta1, double beta2, double eps, double weight_decay, const float *step,
                     void *stream);

/* The same step on grad * (*grad_scale): `grad_scale` is a device float (or NULL = 1) holding the coefficient of
 * torch.nn.utils.clip_grad_norm_ (/root/reference/engine.py:165-166), times 1 / world_size when `grad` holds the
 * rank SUM of an all-reduce - the scaling pass over the gradient buffer is folded into the optimizer's read. */
int rlipv2_adamw_scaled_f32(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, long long n,
                            double lr, double beta1, double beta2, double eps, double weight_decay,
                            const float *step, const float *grad_scale, void *stream);

/* The same step with the learning rate read from device memory (`lr_dev`, or NULL = the host value `lr`): inside a captured
 * CUDA graph a host scalar is frozen at capture time, so the reference's `StepLR(optimizer, lr_drop)` (main.py:554, 724)
 * would have no effect; a device scalar follows `GraphedParSeDATrainStep.set_lr()`.  `skip_flag` (or NULL): a device word
 * that makes the update a no-op when non-zero - the error word of rlipv2_wait_host_flag, so that a replay whose host
 * assignment timed out cannot train on the previous step's indices. */
int rlipv2_adamw_dev_f32(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, long long n, double lr,
                         const float *lr_dev, double beta1, double beta2, double eps, double weight_decay,
                         const float *step, const float *grad_scale, const unsigned *skip_flag, void *stream);

/* Gather scattered fp32 arrays into one flat buffer: for chunk c, copy table[3c+2] elements from the device
 * address table[3c] to dst + table[3c+1] (`table` is a device array of n_chunks x 3 int64; one CTA per chunk,
 * keep chunks <= 64K elements).  Used once per step to collect the gradient tensors autograd produced
 * (wherever it allocated them) into the flat gradient buffer that the all-reduce, the norm and the flat AdamW
 * read - instead of ~600 `grad += new` accumulation kernels into pre-assigned views (main.py:515-517 DDP
 * buckets / engine.py:166-172 in the reference). */
int rlipv2_gather_chunks_f32(const long long *table, int n_chunks, float *dst, void *stream);

/* Head of the backward graph: a one-thread kernel that holds `stream` until the 32-bit word at `flag` (pinned host
 * memory, device-visible) reaches *seq + 1, then stores that value to *seq.  The host publishes the replay's
 * sequence number after it has written the matched indices (models/matcher.py:193 runs scipy on the host) into the
 * pinned buffers that the following copy nodes read; the graph launch itself no longer waits for the host.  After
 * `timeout_ns` without the flag the kernel stores the awaited value to *err and lets the stream continue. */
int rlipv2_wait_host_flag(const unsigned *flag, unsigned *seq, unsigned long long timeout_ns, unsigned *err,
                          void *stream);

/* Diagnostic: a one-thread kernel that stores the GPU's %globaltimer (ns) to *dst when the stream reaches it. */
int rlipv2_stamp_globaltimer(unsigned long long *dst, void *stream);

/* y[i] = sigmoid(delta[i] + inverse_sigmoid(ref[i])): the DAB decoder's box refinement
 * (/root/reference/models/dab_deformable/deformable_transformer.py:1511-1541, util/misc.py:460-464) in one pass. */
int rlipv2_box_refine_f32(const float *delta, const float *ref, float eps, long long n, float *y, void *stream);

/* gen_sineembed_for_position (deformable_transformer.py:1777-1802): pos [rows, n] (n = 2 or 4; x, y[, w, h]) ->
 * out [rows, n*128], coordinate order (y, x[, w, h]), temperature 10000. */
int rlipv2_sine_embed_f32(const float *pos, int rows, int n, float *out, void *stream);

/* Matched-pair box losses of SetCriterionHOI (/root/reference/models/hoi.py:4162-4193; util/box_ops.py:19-73): for
 * `rows` pairs of (cx, cy, w, h) boxes, l1[r] = sum |src - tgt|, giou_loss[r] = 1 - GIoU(src, tgt), and their
 * gradients w.r.t. src: dl1[rows, 4], dgiou[rows, 4] (the reverse-mode derivative torch's autograd forms for the
 * same expression).  src, tgt, dl1, dgiou 16-byte aligned. */
int rlipv2_box_pair_loss_f32(const float *src, const float *tgt, int rows, float *l1, float *giou_loss, float *dl1,
                             float *dgiou, void *stream);

/* nn.GroupNorm(32, 256) of the input projections (/root/reference/models/hoi.py:1937-1952) on token-major activations:
 * x [N, HW, 256] (what an NHWC convolution produces) -> out rows [n * out_batch_stride + hw * 256 ...], i.e. straight
 * into this level's rows of the encoder's [N, sum HW, 256] token buffer (out points at the level's first row,
 * out_batch_stride = sum HW * 256).  stats: scratch of N*32*2 doubles; mean / rstd [N, 32] are kept for the backward.
 * Backward: dy addressed like `out`, dx [N, HW, 256]; dgamma / dbeta [256] are ADDED to (caller zero-fills or passes the
 * running gradient); sums: scratch of N*32*2 doubles.  Only C = 256, G = 32 (returns ESHAPE otherwise). */
int rlipv2_groupnorm_tokens_fwd_f32(const float *x, const float *gamma, const float *beta, float eps, int N, int HW,
                                    int C, int G, double *stats, float *out, long long out_batch_stride, float *mean,
                                    float *rstd, void *stream);
int rlipv2_groupnorm_tokens_bwd_f32(const float *dy, long long dy_batch_stride, const float *x, const float *mean,
                                    const float *rstd, const float *gamma, int N, int HW, int C, int G, double *sums,
                                    float *dx, float *dgamma, float *dbeta, void *stream);

/* Multi-head attention over very short sequences (T <= 8 tokens, head dim 64): the self-attention of the text tower's
 * RoBERTa layers on the label strings (HF transformers modeling_roberta.eager_attention_forward: scores = q k^T * scale
 * + mask, softmax, dropout, probs v; called through /root/reference/models/dab_deformable/deformable_transformer.py:497-502).
 * q, k, v, out, grad_out, dq, dk, dv: [B, T, H, 64] (the layout of the projections); mask: additive [B, T] over the keys
 * or NULL; dropout_p in [0, 1): kept elements are scaled by 1 / (1 - p), the keep decision is a hash of (*seed, salt,
 * element) - `seed` is a device int64, `salt` distinguishes call sites; the forward stores the seed it read in
 * *seed_used (may be NULL) and the backward must be given that value to regenerate the mask. */
int rlipv2_short_attention_fwd_f32(const float *q, const float *k, const float *v, const float *mask, int B, int H, int T,
                                   int D, float scale, double dropout_p, const long long *seed, unsigned salt, float *out,
                                   long long *seed_used, void *stream);
int rlipv2_short_attention_bwd_f32(const float *q, const float *k, const float *v, const float *mask, const float *grad_out,
                                   int B, int H, int T, int D, float scale, double dropout_p, const long long *seed,
                                   unsigned salt, float *dq, float *dk, float *dv, void *stream);

const char *rlipv2_fused_error_string(int code);
unsigned long long rlipv2_fused_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* RLIPV2_FUSED_H_ */
