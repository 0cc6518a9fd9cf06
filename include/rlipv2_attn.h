/* rlipv2_attn.h - C ABI of the fused attention cores of the ParSeDA hot path (tcgen05 / TMEM / TMA, sm_100a).
 *
 * One kernel family serves the three "softmax(Q K^T) V over a few hundred keys" contractions of the step:
 *
 *   ALIF bidirectional cross-attention   /root/reference/models/fuse_helper.py:395-445
 *        `attn_weights = bmm(q, k^T)`; `softmax(attn_weights, -1)`; `softmax(attn_weights^T - rowmax, -1)`; dropout on both;
 *        `bmm(probs_v, value_l)`, `bmm(probs_l, value_v)` - 8 heads x head_dim 256, Tv = 273 image tokens, Tl = 256 labels.
 *        Both directions are softmax-attention over the SAME score matrix read row-wise / column-wise, i.e.
 *        attention(Q = q, K = k, V = value_l) and attention(Q = k, K = q, V = value_v): two calls of this entry point.
 *   RobertaLayer self-attention          models/modeling_roberta.py:185-241   (12 heads x 64, additive -10000 key mask)
 *   decoder query self-attention         models/dab_deformable/deformable_transformer.py:1361, 1383-1390
 *        (nn.MultiheadAttention 8 x 32 over 300 / 150 queries)
 *
 * Layout: q [B, Tq, H*D], k / v [B, Nk, H*D] fp32, head h = columns [h*D, (h+1)*D) - the layout the projections produce
 * and the output projections consume (no head transposes).  Each tensor is described by a base pointer, a row stride `ld`
 * and a batch stride `bs` (in floats), so that column slices of a fused projection can be passed in place.  Base pointers
 * 16-byte aligned, strides multiples of 4 floats, D in {32, 64, 128, 256}.
 *
 * Arithmetic: TF32 tensor-core products (tcgen05.mma kind::tf32), fp32 accumulation in tensor memory, softmax in fp32
 * (expf, true division) - what the reference's pinned torch 1.10 computes for the same bmm / softmax chain with its default
 * allow_tf32 on a tensor-core GPU.  Dropout keeps an element when a splitmix64 hash of (seed, salt, element index) clears
 * the threshold; the backward regenerates the mask from the seed the forward reports.
 *
 * Forward (attn_fwd_kernel): one CTA per (128-query tile, batch x head, output-column slice); S = Q K^T accumulates in
 * tensor memory over a TMA ring of 32-wide head-dim chunks; four softmax warps read S, write the unnormalised (dropped)
 * probabilities back into the same TMEM columns; the P V product takes its A operand straight from tensor memory
 * (tcgen05.mma with a TMEM A operand) and V as stored (MN-major TMA boxes); the epilogue scales rows by 1 / (sum (1 - p)).
 * Constraint: roundup(Nk, 32) + output slice <= 512 TMEM columns (slices of 256 / 128 / 64 / 32 columns are chosen to
 * fit): Nk <= 480.  Larger problems return RLIPV2_ATTN_ESHAPE.
 *
 * Backward: attn_bwd_ds_kernel recomputes a 128 x 128 score tile and dP = dO V^T on the tensor cores and emits
 * dS = P o (dP~ - delta) * scale and the dropped probabilities P~; three batched tcgen05 GEMMs (attn_bgemm_kernel) form
 * dQ = dS K, dK = dS^T Q, dV = P~^T dO with the operands read as stored.
 */
#ifndef RLIPV2_ATTN_H
#define RLIPV2_ATTN_H

#ifdef __cplusplus
extern "C" {
#endif

#define RLIPV2_ATTN_EINVAL  (-1)  /* null pointer / negative size */
#define RLIPV2_ATTN_ESHAPE  (-2)  /* D not in {32,64,128,256}, Nk > 480, strides not multiples of 4 */
#define RLIPV2_ATTN_EALIGN  (-3)  /* base pointer not 16-byte aligned */
#define RLIPV2_ATTN_EDRIVER (-4)  /* cuTensorMapEncodeTiled unavailable or failed */

/* 1 when the forward / backward entry points accept the problem (see the constraints above). */
int rlipv2_attn_supported(int B, int H, int Tq, int Nk, int D);

/* roundup(Nk, 32): the row pitch (in floats) of the backward's two [B*H, Tq, pitch] workspaces. */
int rlipv2_attn_key_pitch(int Nk);

/* out[b, i, h*D + c] = sum_j dropout(softmax_j(scale * <q[b,i,h], k[b,j,h]> + key_bias[b, j])) v[b, j, h*D + c]
 * key_bias: [B, Nk] additive (may be NULL).  stats: [B*H, Tq, 2] = (row maximum, row sum of exp) for the backward.
 * dropout_p in [0, 1); seed: device int64 (read when dropout_p > 0; may be NULL = 0); seed_used: device int64 written with
 * the seed the masks were hashed from (may be NULL). */
int rlipv2_attn_forward_tf32(const float *q, long long q_ld, long long q_bs, const float *k, long long k_ld, long long k_bs,
                             const float *v, long long v_ld, long long v_bs, const float *key_bias, float *out,
                             long long o_ld, long long o_bs, float *stats, int B, int H, int Tq, int Nk, int D, float scale,
                             double dropout_p, const long long *seed, unsigned salt, long long *seed_used, void *stream);

/* Gradients of the call above.  `out` / `dout`: the forward's result and its gradient (same ld / bs as each other);
 * ws_ds, ws_p: workspaces of B*H*Tq*rlipv2_attn_key_pitch(Nk) floats each.  accumulate != 0: dq / dk / dv are added to
 * (red.global.add) instead of stored - the two directions of ALIF sum into the same q / k gradients. */
int rlipv2_attn_backward_tf32(const float *q, long long q_ld, long long q_bs, const float *k, long long k_ld, long long k_bs,
                              const float *v, long long v_ld, long long v_bs, const float *key_bias, const float *out,
                              const float *dout, long long o_ld, long long o_bs, const float *stats, float *dq,
                              long long dq_ld, long long dq_bs, float *dk, long long dk_ld, long long dk_bs, float *dv,
                              long long dv_ld, long long dv_bs, float *ws_ds, float *ws_p, int B, int H, int Tq, int Nk,
                              int D, float scale, double dropout_p, const long long *seed_used, unsigned salt,
                              int accumulate, void *stream);

const char *rlipv2_attn_error_string(int code);

/* kernels launched through this library since load (bench.py's gpu_launches) */
unsigned long long rlipv2_attn_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
