// fused_ops.cu - HBM-bound fused kernels of the ParSeDA train step (sm_100a).  See include/rlipv2_fused.h
// for what each replaces in the reference.  All of them stream their operands exactly once with 128-bit
// coalesced accesses; grids are sized in multiples of the 148 SMs.
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "rlipv2_fused.h"

namespace {

std::atomic<unsigned long long> g_launches{0};
constexpr int kSMs = 148;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

// ---- add + LayerNorm forward: one warp per row, C = 128 * VPL elements (VPL float4 per lane) -------------
template <int VPL>
__global__ void __launch_bounds__(256)
add_layernorm_fwd_kernel(const float *__restrict__ x, const float *__restrict__ r, const float *__restrict__ gamma,
                         const float *__restrict__ beta, float eps, int M, float *__restrict__ y,
                         float *__restrict__ z, float *__restrict__ mean, float *__restrict__ rstd)
{
    constexpr int C = 128 * VPL;
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    float4 gm[VPL], bt[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        gm[i] = __ldg(reinterpret_cast<const float4 *>(gamma) + i * 32 + lane);
        bt[i] = __ldg(reinterpret_cast<const float4 *>(beta) + i * 32 + lane);
    }
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < M; row += warps) {
        const float4 *xr = reinterpret_cast<const float4 *>(x + (size_t)row * C);
        const float4 *rr = r ? reinterpret_cast<const float4 *>(r + (size_t)row * C) : nullptr;
        float4 v[VPL];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            v[i] = xr[i * 32 + lane];
            if (rr) { const float4 t = rr[i * 32 + lane]; v[i].x += t.x; v[i].y += t.y; v[i].z += t.z; v[i].w += t.w; }
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
        const float mu = warp_sum(s) * (1.f / C);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
            q += (a * a + b * b) + (c * c + d * d);
        }
        const float rs = rsqrtf(warp_sum(q) * (1.f / C) + eps);
        float4 *yr = reinterpret_cast<float4 *>(y + (size_t)row * C);
        float4 *zr = reinterpret_cast<float4 *>(z + (size_t)row * C);
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            zr[i * 32 + lane] = v[i];
            float4 o;
            o.x = (v[i].x - mu) * rs * gm[i].x + bt[i].x;
            o.y = (v[i].y - mu) * rs * gm[i].y + bt[i].y;
            o.z = (v[i].z - mu) * rs * gm[i].z + bt[i].z;
            o.w = (v[i].w - mu) * rs * gm[i].w + bt[i].w;
            yr[i * 32 + lane] = o;
        }
        if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
    }
}

// ---- LayerNorm backward: dz per row + dgamma/dbeta accumulated per lane over the warp's rows, reduced
// over the CTA's 8 warps in shared memory, one atomicAdd per (CTA, column) ------------------------------------
template <int VPL>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float *__restrict__ dy, const float *__restrict__ z, const float *__restrict__ mean,
                     const float *__restrict__ rstd, const float *__restrict__ gamma, int M,
                     float *__restrict__ dz, float *__restrict__ dgamma, float *__restrict__ dbeta)
{
    constexpr int C = 128 * VPL;
    __shared__ float red[8][C];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    float4 gm[VPL], ag[VPL], ab[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        gm[i] = __ldg(reinterpret_cast<const float4 *>(gamma) + i * 32 + lane);
        ag[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < M; row += warps) {
        const float4 *dr = reinterpret_cast<const float4 *>(dy + (size_t)row * C);
        const float4 *zr = reinterpret_cast<const float4 *>(z + (size_t)row * C);
        const float mu = mean[row], rs = rstd[row];
        float4 g[VPL], xh[VPL];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const float4 d = dr[i * 32 + lane], v = zr[i * 32 + lane];
            xh[i] = make_float4((v.x - mu) * rs, (v.y - mu) * rs, (v.z - mu) * rs, (v.w - mu) * rs);
            g[i] = make_float4(d.x * gm[i].x, d.y * gm[i].y, d.z * gm[i].z, d.w * gm[i].w);
            s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
            s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
            ag[i].x += d.x * xh[i].x; ag[i].y += d.y * xh[i].y; ag[i].z += d.z * xh[i].z; ag[i].w += d.w * xh[i].w;
            ab[i].x += d.x; ab[i].y += d.y; ab[i].z += d.z; ab[i].w += d.w;
        }
        const float c1 = warp_sum(s1) * (1.f / C), c2 = warp_sum(s2) * (1.f / C);
        float4 *or_ = reinterpret_cast<float4 *>(dz + (size_t)row * C);
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            float4 o;
            o.x = rs * (g[i].x - c1 - xh[i].x * c2);
            o.y = rs * (g[i].y - c1 - xh[i].y * c2);
            o.z = rs * (g[i].z - c1 - xh[i].z * c2);
            o.w = rs * (g[i].w - c1 - xh[i].w * c2);
            or_[i * 32 + lane] = o;
        }
    }
    // CTA reduction of the per-warp column partials, then one atomic per column
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int i = 0; i < VPL; ++i)
            *reinterpret_cast<float4 *>(&red[warp][(i * 32 + lane) * 4]) = pass ? ab[i] : ag[i];
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += red[w][c];
            atomicAdd((pass ? dbeta : dgamma) + c, s);
        }
        __syncthreads();
    }
}

// ---- (ReLU mask +) column sum: CTA = 32 columns x 8 row-lanes, grid.y row chunks --------------------------------
// MASK: 0 = plain column sum, 1 = ReLU mask from the forward output y, 2 = row mask (bytes, y reinterpreted):
// rows with a non-zero byte are zeroed (the backward of masked_fill(mask[..., None], 0))
template <int MASK>
__global__ void __launch_bounds__(256)
relu_bwd_colsum_kernel(const float *__restrict__ g, const float *__restrict__ y, float *__restrict__ gm,
                       float *__restrict__ colsum, int M, int N)
{
    // each thread owns 4 consecutive columns (float4): 8 lanes cover 32 columns = 128 bytes of a row
    __shared__ float4 red[32][8];
    const int cl = threadIdx.x & 7;              // float4 index inside the 32-column tile
    const int rl = threadIdx.x >> 3;             // 0..31 row lane
    const int col = blockIdx.x * 32 + cl * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int row = blockIdx.y * 32 + rl; row < M; row += gridDim.y * 32) {
        const size_t off = (size_t)row * N + col;
        float4 v = *reinterpret_cast<const float4 *>(g + off);
        if (MASK == 1) {
            const float4 t = *reinterpret_cast<const float4 *>(y + off);
            v.x = t.x > 0.f ? v.x : 0.f; v.y = t.y > 0.f ? v.y : 0.f;
            v.z = t.z > 0.f ? v.z : 0.f; v.w = t.w > 0.f ? v.w : 0.f;
            *reinterpret_cast<float4 *>(gm + off) = v;
        } else if (MASK == 2) {
            if (reinterpret_cast<const unsigned char *>(y)[row]) v = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4 *>(gm + off) = v;
        }
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    red[rl][cl] = acc;
    __syncthreads();
    if (rl == 0) {
        float4 s = red[0][cl];
        for (int r = 1; r < 32; ++r) { const float4 t = red[r][cl]; s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
        atomicAdd(colsum + col, s.x); atomicAdd(colsum + col + 1, s.y);
        atomicAdd(colsum + col + 2, s.z); atomicAdd(colsum + col + 3, s.w);
    }
}

// ---- flat AdamW -------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adamw_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
             long long n4, long long n, float lr, float b1, float b2, float omb1, float omb2, float eps, float wd,
             const float *__restrict__ step, const float *__restrict__ grad_scale, const float *__restrict__ lr_dev,
             const unsigned *__restrict__ skip_flag)
{
    // skip_flag: the error word of the backward graph's host-flag wait - a replay whose assignment never arrived must
    // not update the parameters from the previous step's indices (the update becomes a no-op; check() raises later)
    if (skip_flag && *skip_flag != 0u) return;
    // lr_dev: the learning rate as a device scalar, so that a captured graph follows the scheduler (main.py:554 StepLR)
    if (lr_dev) lr = *lr_dev;
    // omb1 / omb2 = 1 - beta computed in double on the host (1.f - 0.999f is 1.3e-5 off)
    const float t = *step;
    // grad_scale: the clip_grad_norm_ coefficient (and 1 / world for the rank mean) applied to the gradient as it is
    // read, instead of a separate pass over the gradient buffer
    const float gs = grad_scale ? *grad_scale : 1.f;
    const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
    const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2), decay = 1.f - lr * wd;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4 *>(p)[i], mm = reinterpret_cast<float4 *>(m)[i], vv = reinterpret_cast<float4 *>(v)[i];
        float4 gg = reinterpret_cast<const float4 *>(g)[i];
        gg.x *= gs; gg.y *= gs; gg.z *= gs; gg.w *= gs;
#define RLIPV2_ADAMW_LANE(c)                                                         \
        pp.c *= decay;                                                               \
        mm.c = b1 * mm.c + omb1 * gg.c;                                              \
        vv.c = b2 * vv.c + omb2 * gg.c * gg.c;                                       \
        pp.c -= step_size * mm.c / (sqrtf(vv.c) * inv_sqrt_bc2 + eps);
        RLIPV2_ADAMW_LANE(x) RLIPV2_ADAMW_LANE(y) RLIPV2_ADAMW_LANE(z) RLIPV2_ADAMW_LANE(w)
#undef RLIPV2_ADAMW_LANE
        reinterpret_cast<float4 *>(p)[i] = pp;
        reinterpret_cast<float4 *>(m)[i] = mm;
        reinterpret_cast<float4 *>(v)[i] = vv;
    }
    // tail (n not a multiple of 4)
    for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float pp = p[i] * decay;
        const float gg = g[i] * gs;
        const float mm = b1 * m[i] + omb1 * gg, vv = b2 * v[i] + omb2 * gg * gg;
        pp -= step_size * mm / (sqrtf(vv) * inv_sqrt_bc2 + eps);
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
}

// ---- gather of scattered gradient tensors into the flat gradient buffer ------------------------------------
// table[3*c + {0,1,2}] = (source address, destination element offset, element count) of chunk c; one CTA per
// chunk.  128-bit copies when both sides are 16-byte aligned, scalar otherwise.
__global__ void __launch_bounds__(256)
gather_chunks_kernel(const long long *__restrict__ table, float *__restrict__ dst)
{
    const long long *row = table + 3ll * blockIdx.x;
    const float *src = reinterpret_cast<const float *>(row[0]);
    float *d = dst + row[1];
    const int n = (int)row[2];
    if ((((uintptr_t)src | (uintptr_t)d) & 15) == 0) {
        const int n4 = n >> 2;
        for (int i = threadIdx.x; i < n4; i += 256)
            reinterpret_cast<float4 *>(d)[i] = __ldg(reinterpret_cast<const float4 *>(src) + i);
        for (int i = (n4 << 2) + threadIdx.x; i < n; i += 256) d[i] = src[i];
    } else {
        for (int i = threadIdx.x; i < n; i += 256) d[i] = src[i];
    }
}

// ---- host-flag wait: the head of the backward graph ------------------------------------------------------------
// The step's only host work is the assignment solve between the forward graph and the backward graph.  The
// backward graph is launched right behind the forward graph (its multi-millisecond launch cost overlaps the
// forward's execution) and starts with this one-thread kernel, which holds the stream until the host publishes
// the sequence number of this replay in a pinned, device-visible word - after it has written the matched indices
// into the pinned buffers the following memcpy nodes read.  A timeout (reported through *err) bounds the wait, so a
// host that died cannot hang the GPU.
__global__ void wait_host_flag_kernel(const unsigned *flag, unsigned *seq, unsigned long long timeout_ns, unsigned *err)
{
    const unsigned want = *seq + 1u;
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int)(v - want) >= 0) break;
        __nanosleep(200);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > timeout_ns) { *err = want; break; }
    }
    *seq = want;
}

// one-thread timestamp: *dst = %globaltimer (ns).  Diagnostic node for the graphed step's anatomy (tools/step_anatomy.py).
__global__ void stamp_kernel(unsigned long long *dst)
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    *dst = t;
}

// ---- DAB decoder small-op chains -------------------------------------------------------------------------------
// y = sigmoid(delta + inverse_sigmoid(ref)), inverse_sigmoid(x) = log(max(clamp(x,0,1), eps) / max(1 - clamp(x,0,1), eps))
// (util/misc.py:460-464 + the box refinement of dab_deformable/deformable_transformer.py:1511-1541): 9 torch
// kernels over a [bs, nq, 4] tensor, one here.  Same operation order and IEEE division / accurate logf, expf.
__global__ void __launch_bounds__(256)
box_refine_kernel(const float *__restrict__ delta, const float *__restrict__ ref, float eps, long long n,
                  float *__restrict__ y)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float x = fminf(fmaxf(ref[i], 0.f), 1.f);
        const float inv = logf(__fdiv_rn(fmaxf(x, eps), fmaxf(1.f - x, eps)));
        const float z = delta[i] + inv;
        y[i] = __fdiv_rn(1.f, 1.f + expf(-z));
    }
}

// Matched-pair box losses of SetCriterionHOI (/root/reference/models/hoi.py:4162-4193 with util/box_ops.py:19-73):
// per row r, src/tgt boxes (cx, cy, w, h):
//   l1[r]   = sum_k |src_k - tgt_k|                    gl[r] = 1 - GIoU(xyxy(src), xyxy(tgt))
// and their gradients w.r.t. src, dl1[r, 4] = sign(src - tgt), dgl[r, 4] - the reverse-mode derivative of the same
// expression graph torch differentiates (clamp(min=0) passes gradient where its argument is >= 0, min / max split
// ties evenly).  The forward values use the operation order of the torch formulation with explicitly rounded
// products / sums (no FMA contraction), so they equal torch's bit for bit.  One thread per row: the reference spends
// ~250 kernel launches (forward + backward) on these ~30 rows.
__device__ __forceinline__ float sel_w(float a, float b) { return a > b ? 1.f : (a == b ? 0.5f : 0.f); }

__global__ void __launch_bounds__(128)
box_pair_loss_kernel(const float *__restrict__ src, const float *__restrict__ tgt, int R, float *__restrict__ l1,
                     float *__restrict__ gl, float *__restrict__ dl1, float *__restrict__ dgl)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const float4 s = reinterpret_cast<const float4 *>(src)[r], t = reinterpret_cast<const float4 *>(tgt)[r];
    // box_cxcywh_to_xyxy
    const float a0x = __fsub_rn(s.x, __fmul_rn(0.5f, s.z)), a0y = __fsub_rn(s.y, __fmul_rn(0.5f, s.w));
    const float a1x = __fadd_rn(s.x, __fmul_rn(0.5f, s.z)), a1y = __fadd_rn(s.y, __fmul_rn(0.5f, s.w));
    const float b0x = __fsub_rn(t.x, __fmul_rn(0.5f, t.z)), b0y = __fsub_rn(t.y, __fmul_rn(0.5f, t.w));
    const float b1x = __fadd_rn(t.x, __fmul_rn(0.5f, t.z)), b1y = __fadd_rn(t.y, __fmul_rn(0.5f, t.w));
    const float aw = __fsub_rn(a1x, a0x), ah = __fsub_rn(a1y, a0y);
    const float area_a = __fmul_rn(aw, ah), area_b = __fmul_rn(__fsub_rn(b1x, b0x), __fsub_rn(b1y, b0y));
    const float dix = __fsub_rn(fminf(a1x, b1x), fmaxf(a0x, b0x)), diy = __fsub_rn(fminf(a1y, b1y), fmaxf(a0y, b0y));
    const float iw = fmaxf(dix, 0.f), ih = fmaxf(diy, 0.f);
    const float inter = __fmul_rn(iw, ih);
    const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    const float iou = __fdiv_rn(inter, uni);
    const float dex = __fsub_rn(fmaxf(a1x, b1x), fminf(a0x, b0x)), dey = __fsub_rn(fmaxf(a1y, b1y), fminf(a0y, b0y));
    const float ew = fmaxf(dex, 0.f), eh = fmaxf(dey, 0.f);
    const float earea = __fmul_rn(ew, eh);
    const float giou = __fsub_rn(iou, __fdiv_rn(__fsub_rn(earea, uni), earea));
    gl[r] = __fsub_rn(1.f, giou);
    const float e0 = __fsub_rn(s.x, t.x), e1 = __fsub_rn(s.y, t.y), e2 = __fsub_rn(s.z, t.z), e3 = __fsub_rn(s.w, t.w);
    l1[r] = __fadd_rn(__fadd_rn(__fadd_rn(fabsf(e0), fabsf(e1)), fabsf(e2)), fabsf(e3));
    auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
    reinterpret_cast<float4 *>(dl1)[r] = make_float4(sgn(e0), sgn(e1), sgn(e2), sgn(e3));
    // d giou / d (a0x, a0y, a1x, a1y)
    const float G_union = -inter / (uni * uni) + 1.f / earea;
    const float g_inter = 1.f / uni - G_union;
    const float g_earea = -uni / (earea * earea);
    const float g_iw = dix >= 0.f ? g_inter * ih : 0.f, g_ih = diy >= 0.f ? g_inter * iw : 0.f;
    const float g_ew = dex >= 0.f ? g_earea * eh : 0.f, g_eh = dey >= 0.f ? g_earea * ew : 0.f;
    const float g_a1x = g_iw * sel_w(b1x, a1x) + g_ew * sel_w(a1x, b1x) + G_union * ah;
    const float g_a0x = -g_iw * sel_w(a0x, b0x) - g_ew * sel_w(b0x, a0x) - G_union * ah;
    const float g_a1y = g_ih * sel_w(b1y, a1y) + g_eh * sel_w(a1y, b1y) + G_union * aw;
    const float g_a0y = -g_ih * sel_w(a0y, b0y) - g_eh * sel_w(b0y, a0y) - G_union * aw;
    // back through cxcywh -> xyxy, and the sign of (1 - giou)
    reinterpret_cast<float4 *>(dgl)[r] = make_float4(-(g_a0x + g_a1x), -(g_a0y + g_a1y), -0.5f * (g_a1x - g_a0x),
                                                     -0.5f * (g_a1y - g_a0y));
}

// ---- attention over very short sequences (the text tower's label strings: 3-8 tokens) ---------------------------
// HF RobertaSelfAttention's eager path (transformers modeling_roberta.eager_attention_forward, called by the reference at
// /root/reference/models/dab_deformable/deformable_transformer.py:497-502 for every label string, every step) issues
// bmm + scale + mask add + softmax + dropout + bmm + transpose copy per layer - 7 launches forward, ~10 backward, for
// 256 x 12 problems of 5 x 5 scores; the batched-GEMM kernels alone take 37 + 9 us.  Here one warp owns one
// (label, head): q, k, v [T <= 8, 64] go to shared memory, lane (i, j) forms score (i, j), lane i the softmax of row i,
// and every lane two of the 64 output channels.  Inputs / outputs are addressed as [B, T, H, 64] (the layout the
// q/k/v projections produce and the output projection consumes): no transposes.  Dropout keeps an element when a
// splitmix64 hash of (seed, salt, element index) clears the threshold; the backward regenerates the same mask.
constexpr int kSaT = 8, kSaD = 64, kSaWarps = 4;

__device__ __forceinline__ bool sa_keep(unsigned long long seed, unsigned salt, unsigned idx, unsigned thresh) {
    unsigned long long z = seed * 0x9E3779B97F4A7C15ull + ((unsigned long long)salt << 32 | idx) + 0xD1B54A32D192ED03ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (unsigned)(z >> 32) >= thresh;
}

struct SaSmem {
    float q[kSaT][kSaD], k[kSaT][kSaD], v[kSaT][kSaD], g[kSaT][kSaD];
    float p[kSaT][kSaT], pd[kSaT][kSaT], ds[kSaT][kSaT];
};

// scores -> softmax probabilities p and dropped / rescaled probabilities pd, both in shared memory
__device__ __forceinline__ void sa_probs(SaSmem &sm, const float *__restrict__ mask, int b, int h, int H, int T, int lane,
                                         float scale, unsigned thresh, float inv_keep, unsigned long long seed,
                                         unsigned salt)
{
    for (int e = lane; e < T * T; e += 32) {
        const int i = e / T, j = e - i * T;
        float acc = 0.f;
#pragma unroll 16
        for (int d = 0; d < kSaD; ++d) acc = fmaf(sm.q[i][(d + lane) & 63], sm.k[j][(d + lane) & 63], acc);
        sm.p[i][j] = acc * scale + (mask ? mask[(size_t)b * T + j] : 0.f);
    }
    __syncwarp();
    if (lane < T) {
        const int i = lane;
        float mx = -INFINITY;
        for (int j = 0; j < T; ++j) mx = fmaxf(mx, sm.p[i][j]);
        float e[kSaT], sum = 0.f;
        for (int j = 0; j < T; ++j) { e[j] = expf(sm.p[i][j] - mx); sum += e[j]; }
        for (int j = 0; j < T; ++j) {
            const float pr = __fdiv_rn(e[j], sum);
            sm.p[i][j] = pr;
            bool keep = true;
            if (thresh) keep = sa_keep(seed, salt, (unsigned)(((b * H + h) * kSaT + i) * kSaT + j), thresh);
            sm.pd[i][j] = keep ? pr * inv_keep : 0.f;
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kSaWarps * 32)
short_attn_fwd_kernel(const float *__restrict__ q, const float *__restrict__ k, const float *__restrict__ v,
                      const float *__restrict__ mask, int B, int H, int T, float scale, unsigned thresh, float inv_keep,
                      const long long *__restrict__ seed_ptr, unsigned salt, float *__restrict__ out,
                      long long *__restrict__ seed_used)
{
    __shared__ SaSmem smem[kSaWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.x * kSaWarps + warp;
    const unsigned long long seed = seed_ptr ? (unsigned long long)*seed_ptr : 0ull;
    if (blockIdx.x == 0 && threadIdx.x == 0 && seed_used) *seed_used = (long long)seed;
    if (bh >= B * H) return;
    const int b = bh / H, h = bh - b * H;
    SaSmem &sm = smem[warp];
    for (int t = 0; t < T; ++t) {
        const size_t base = (((size_t)b * T + t) * H + h) * kSaD + lane * 2;
        *reinterpret_cast<float2 *>(&sm.q[t][lane * 2]) = *reinterpret_cast<const float2 *>(q + base);
        *reinterpret_cast<float2 *>(&sm.k[t][lane * 2]) = *reinterpret_cast<const float2 *>(k + base);
        *reinterpret_cast<float2 *>(&sm.v[t][lane * 2]) = *reinterpret_cast<const float2 *>(v + base);
    }
    __syncwarp();
    sa_probs(sm, mask, b, h, H, T, lane, scale, thresh, inv_keep, seed, salt);
    for (int i = 0; i < T; ++i) {
        float2 o = make_float2(0.f, 0.f);
        for (int j = 0; j < T; ++j) {
            const float w = sm.pd[i][j];
            o.x = fmaf(w, sm.v[j][lane * 2], o.x);
            o.y = fmaf(w, sm.v[j][lane * 2 + 1], o.y);
        }
        *reinterpret_cast<float2 *>(out + (((size_t)b * T + i) * H + h) * kSaD + lane * 2) = o;
    }
}

__global__ void __launch_bounds__(kSaWarps * 32)
short_attn_bwd_kernel(const float *__restrict__ q, const float *__restrict__ k, const float *__restrict__ v,
                      const float *__restrict__ mask, const float *__restrict__ gout, int B, int H, int T, float scale,
                      unsigned thresh, float inv_keep, const long long *__restrict__ seed_ptr, unsigned salt,
                      float *__restrict__ dq, float *__restrict__ dk, float *__restrict__ dv)
{
    __shared__ SaSmem smem[kSaWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.x * kSaWarps + warp;
    if (bh >= B * H) return;
    const unsigned long long seed = seed_ptr ? (unsigned long long)*seed_ptr : 0ull;
    const int b = bh / H, h = bh - b * H;
    SaSmem &sm = smem[warp];
    for (int t = 0; t < T; ++t) {
        const size_t base = (((size_t)b * T + t) * H + h) * kSaD + lane * 2;
        *reinterpret_cast<float2 *>(&sm.q[t][lane * 2]) = *reinterpret_cast<const float2 *>(q + base);
        *reinterpret_cast<float2 *>(&sm.k[t][lane * 2]) = *reinterpret_cast<const float2 *>(k + base);
        *reinterpret_cast<float2 *>(&sm.v[t][lane * 2]) = *reinterpret_cast<const float2 *>(v + base);
        *reinterpret_cast<float2 *>(&sm.g[t][lane * 2]) = *reinterpret_cast<const float2 *>(gout + base);
    }
    __syncwarp();
    sa_probs(sm, mask, b, h, H, T, lane, scale, thresh, inv_keep, seed, salt);
    // d pd (i, j) = <gout_i, v_j>; through the dropout: d p = d pd * (pd / p) (0 where dropped)
    for (int e = lane; e < T * T; e += 32) {
        const int i = e / T, j = e - i * T;
        float acc = 0.f;
#pragma unroll 16
        for (int d = 0; d < kSaD; ++d) acc = fmaf(sm.g[i][(d + lane) & 63], sm.v[j][(d + lane) & 63], acc);
        sm.ds[i][j] = sm.pd[i][j] != 0.f ? acc * inv_keep : 0.f;          // = d p (i, j)
    }
    __syncwarp();
    if (lane < T) {                                                        // softmax backward of row `lane`
        const int i = lane;
        float dot = 0.f;
        for (int j = 0; j < T; ++j) dot = fmaf(sm.ds[i][j], sm.p[i][j], dot);
        for (int j = 0; j < T; ++j) sm.ds[i][j] = sm.p[i][j] * (sm.ds[i][j] - dot);
    }
    __syncwarp();
    for (int t = 0; t < T; ++t) {
        float2 aq = make_float2(0.f, 0.f), ak = make_float2(0.f, 0.f), av = make_float2(0.f, 0.f);
        for (int u = 0; u < T; ++u) {
            const float s_tu = sm.ds[t][u], s_ut = sm.ds[u][t], p_ut = sm.pd[u][t];
            aq.x = fmaf(s_tu, sm.k[u][lane * 2], aq.x); aq.y = fmaf(s_tu, sm.k[u][lane * 2 + 1], aq.y);
            ak.x = fmaf(s_ut, sm.q[u][lane * 2], ak.x); ak.y = fmaf(s_ut, sm.q[u][lane * 2 + 1], ak.y);
            av.x = fmaf(p_ut, sm.g[u][lane * 2], av.x); av.y = fmaf(p_ut, sm.g[u][lane * 2 + 1], av.y);
        }
        const size_t base = (((size_t)b * T + t) * H + h) * kSaD + lane * 2;
        *reinterpret_cast<float2 *>(dq + base) = make_float2(aq.x * scale, aq.y * scale);
        *reinterpret_cast<float2 *>(dk + base) = make_float2(ak.x * scale, ak.y * scale);
        *reinterpret_cast<float2 *>(dv + base) = av;
    }
}

// ---- GroupNorm(32 groups, 256 channels) on token-major activations ------------------------------------------
// The input projections of RLIP_ParSeDA (/root/reference/models/hoi.py:1937-1952: 1x1 / 3x3 conv + nn.GroupNorm(32, 256)
// per feature level) feed the encoder, which wants tokens: [N, sum_l H_l W_l, 256].  cuDNN's TF32 convolutions produce
// NHWC (= token-major) outputs; torch's GroupNorm converts them to NCHW, normalises, and flatten + transpose + cat
// convert back.  These kernels normalise the token-major tensor in place of all that and write straight into the
// level's rows of the concatenated buffer.  Thread t of a 256-thread CTA owns channels 4*(t & 63) .. +3 (a group =
// 8 channels = two neighbouring lanes) of rows (t >> 6) + 4k of the CTA's 64-row chunk.
constexpr int kGnC = 256, kGnG = 32, kGnRows = 64;

__global__ void __launch_bounds__(256)
gn_tok_stats_kernel(const float *__restrict__ x, int HW, double *__restrict__ stats /* [N, 32, 2] */)
{
    const int n = blockIdx.y, c4 = threadIdx.x & 63, r0 = threadIdx.x >> 6;
    const int row_end = min(HW, (int)(blockIdx.x + 1) * kGnRows);
    const float4 *xp = reinterpret_cast<const float4 *>(x + (size_t)n * HW * kGnC);
    float s = 0.f, ss = 0.f;
    for (int r = blockIdx.x * kGnRows + r0; r < row_end; r += 4) {
        const float4 v = xp[(size_t)r * 64 + c4];
        s += (v.x + v.y) + (v.z + v.w);
        ss = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, ss))));
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    __shared__ float part[4][kGnG][2];
    if ((c4 & 1) == 0) { part[r0][c4 >> 1][0] = s; part[r0][c4 >> 1][1] = ss; }
    __syncthreads();
    if (threadIdx.x < 2 * kGnG) {
        const int g = threadIdx.x >> 1, k = threadIdx.x & 1;
        const double v = (double)part[0][g][k] + (double)part[1][g][k] + (double)part[2][g][k] + (double)part[3][g][k];
        atomicAdd(stats + ((size_t)n * kGnG + g) * 2 + k, v);
    }
}

__global__ void __launch_bounds__(256)
gn_tok_apply_kernel(const float *__restrict__ x, const double *__restrict__ stats, const float *__restrict__ gamma,
                    const float *__restrict__ beta, float eps, int N, int HW, float *__restrict__ out,
                    long long out_batch_stride, float *__restrict__ mean, float *__restrict__ rstd)
{
    const long long total = (long long)N * HW * 64;
    const double inv_m = 1.0 / ((double)HW * 8.0);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i & 63);
        const long long row = i >> 6;                    // n * HW + hw
        const int n = (int)(row / HW), hw = (int)(row - (long long)n * HW), g = c4 >> 1;
        const double mu = stats[((size_t)n * kGnG + g) * 2] * inv_m;
        const double var = stats[((size_t)n * kGnG + g) * 2 + 1] * inv_m - mu * mu;
        const float m = (float)mu, rs = rsqrtf((float)(var > 0.0 ? var : 0.0) + eps);
        if (hw == 0 && (c4 & 1) == 0) { mean[n * kGnG + g] = m; rstd[n * kGnG + g] = rs; }
        const float4 v = reinterpret_cast<const float4 *>(x)[i];
        const float4 ga = __ldg(reinterpret_cast<const float4 *>(gamma) + c4);
        const float4 be = __ldg(reinterpret_cast<const float4 *>(beta) + c4);
        float4 o;
        o.x = fmaf((v.x - m) * rs, ga.x, be.x); o.y = fmaf((v.y - m) * rs, ga.y, be.y);
        o.z = fmaf((v.z - m) * rs, ga.z, be.z); o.w = fmaf((v.w - m) * rs, ga.w, be.w);
        *reinterpret_cast<float4 *>(out + (size_t)n * out_batch_stride + (size_t)hw * kGnC + c4 * 4) = o;
    }
}

// backward, pass 1: per (n, group) A = sum dy gamma xhat, B = sum dy gamma; per channel dgamma += sum dy xhat,
// dbeta += sum dy (added into the given buffers)
__global__ void __launch_bounds__(256)
gn_tok_bwd_stats_kernel(const float *__restrict__ dy, long long dy_batch_stride, const float *__restrict__ x,
                        const float *__restrict__ mean, const float *__restrict__ rstd, const float *__restrict__ gamma,
                        int HW, double *__restrict__ sums /* [N, 32, 2] */, float *__restrict__ dgamma,
                        float *__restrict__ dbeta)
{
    const int n = blockIdx.y, c4 = threadIdx.x & 63, r0 = threadIdx.x >> 6, g = c4 >> 1;
    const int row_end = min(HW, (int)(blockIdx.x + 1) * kGnRows);
    const float m = mean[n * kGnG + g], rs = rstd[n * kGnG + g];
    const float4 ga = __ldg(reinterpret_cast<const float4 *>(gamma) + c4);
    const float4 *xp = reinterpret_cast<const float4 *>(x + (size_t)n * HW * kGnC);
    const float *dp = dy + (size_t)n * dy_batch_stride;
    float4 dg = make_float4(0.f, 0.f, 0.f, 0.f), db = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = blockIdx.x * kGnRows + r0; r < row_end; r += 4) {
        const float4 v = xp[(size_t)r * 64 + c4];
        const float4 d = *reinterpret_cast<const float4 *>(dp + (size_t)r * kGnC + c4 * 4);
        dg.x = fmaf(d.x, (v.x - m) * rs, dg.x); dg.y = fmaf(d.y, (v.y - m) * rs, dg.y);
        dg.z = fmaf(d.z, (v.z - m) * rs, dg.z); dg.w = fmaf(d.w, (v.w - m) * rs, dg.w);
        db.x += d.x; db.y += d.y; db.z += d.z; db.w += d.w;
    }
    float A = fmaf(ga.x, dg.x, fmaf(ga.y, dg.y, fmaf(ga.z, dg.z, ga.w * dg.w)));
    float B = fmaf(ga.x, db.x, fmaf(ga.y, db.y, fmaf(ga.z, db.z, ga.w * db.w)));
    A += __shfl_xor_sync(0xffffffffu, A, 1);
    B += __shfl_xor_sync(0xffffffffu, B, 1);
    __shared__ float part[4][kGnG][2];
    __shared__ float4 pg[4][64], pb[4][64];
    if ((c4 & 1) == 0) { part[r0][g][0] = A; part[r0][g][1] = B; }
    pg[r0][c4] = dg;
    pb[r0][c4] = db;
    __syncthreads();
    if (threadIdx.x < 2 * kGnG) {
        const int gg = threadIdx.x >> 1, k = threadIdx.x & 1;
        const double v = (double)part[0][gg][k] + (double)part[1][gg][k] + (double)part[2][gg][k] + (double)part[3][gg][k];
        atomicAdd(sums + ((size_t)n * kGnG + gg) * 2 + k, v);
    }
    if (threadIdx.x >= 64 && threadIdx.x < 128) {
        const int c = threadIdx.x - 64;
        const float4 a = pg[0][c], b = pg[1][c], cc = pg[2][c], d = pg[3][c];
        atomicAdd(dgamma + c * 4 + 0, (a.x + b.x) + (cc.x + d.x)); atomicAdd(dgamma + c * 4 + 1, (a.y + b.y) + (cc.y + d.y));
        atomicAdd(dgamma + c * 4 + 2, (a.z + b.z) + (cc.z + d.z)); atomicAdd(dgamma + c * 4 + 3, (a.w + b.w) + (cc.w + d.w));
    } else if (threadIdx.x >= 128 && threadIdx.x < 192) {
        const int c = threadIdx.x - 128;
        const float4 a = pb[0][c], b = pb[1][c], cc = pb[2][c], d = pb[3][c];
        atomicAdd(dbeta + c * 4 + 0, (a.x + b.x) + (cc.x + d.x)); atomicAdd(dbeta + c * 4 + 1, (a.y + b.y) + (cc.y + d.y));
        atomicAdd(dbeta + c * 4 + 2, (a.z + b.z) + (cc.z + d.z)); atomicAdd(dbeta + c * 4 + 3, (a.w + b.w) + (cc.w + d.w));
    }
}

// backward, pass 2: dx = rstd * (dy gamma - (B + xhat A) / m)
__global__ void __launch_bounds__(256)
gn_tok_bwd_apply_kernel(const float *__restrict__ dy, long long dy_batch_stride, const float *__restrict__ x,
                        const float *__restrict__ mean, const float *__restrict__ rstd, const float *__restrict__ gamma,
                        const double *__restrict__ sums, int N, int HW, float *__restrict__ dx)
{
    const long long total = (long long)N * HW * 64;
    const float inv_m = 1.f / ((float)HW * 8.f);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i & 63);
        const long long row = i >> 6;
        const int n = (int)(row / HW), hw = (int)(row - (long long)n * HW), g = c4 >> 1;
        const float m = mean[n * kGnG + g], rs = rstd[n * kGnG + g];
        const float A = (float)sums[((size_t)n * kGnG + g) * 2] * inv_m, B = (float)sums[((size_t)n * kGnG + g) * 2 + 1] * inv_m;
        const float4 v = reinterpret_cast<const float4 *>(x)[i];
        const float4 d = *reinterpret_cast<const float4 *>(dy + (size_t)n * dy_batch_stride + (size_t)hw * kGnC + c4 * 4);
        const float4 ga = __ldg(reinterpret_cast<const float4 *>(gamma) + c4);
        float4 o;
        o.x = rs * (d.x * ga.x - fmaf((v.x - m) * rs, A, B)); o.y = rs * (d.y * ga.y - fmaf((v.y - m) * rs, A, B));
        o.z = rs * (d.z * ga.z - fmaf((v.z - m) * rs, A, B)); o.w = rs * (d.w * ga.w - fmaf((v.w - m) * rs, A, B));
        reinterpret_cast<float4 *>(dx)[i] = o;
    }
}

// Sine embedding of anchor coordinates (gen_sineembed_for_position, deformable_transformer.py:1777-1802):
// pos [R, n] (x, y[, w, h]) -> out [R, n*128] in the order (y, x[, w, h]); feature k of a coordinate p is
// sin(2 pi p / T_k) for even k, cos(2 pi p / T_k) for odd k, T_k = 10000^(2 floor(k/2) / 128).
__global__ void __launch_bounds__(128)
sine_embed_kernel(const float *__restrict__ pos, int R, int n, float *__restrict__ out)
{
    const int k = threadIdx.x;                       // feature 0..127
    const float dim_t = powf(10000.f, (float)(2 * (k >> 1)) / 128.f);
    const float scale = 6.283185307179586f;
    for (int rc = blockIdx.x; rc < R * n; rc += gridDim.x) {
        const int r = rc / n, c = rc - r * n;
        const int src = c == 0 ? 1 : (c == 1 ? 0 : c);           // (y, x, w, h) <- columns (1, 0, 2, 3)
        const float p = __fdiv_rn(pos[(size_t)r * n + src] * scale, dim_t);
        out[(size_t)rc * 128 + k] = (k & 1) ? cosf(p) : sinf(p);
    }
}

inline int done() {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

}  // namespace

extern "C" {

int rlipv2_add_layernorm_fwd_f32(const float *x, const float *r, const float *gamma, const float *beta, float eps,
                                 int M, int C, float *y, float *z, float *mean, float *rstd, void *stream)
{
    if (M == 0) return 0;
    if (!x || !gamma || !beta || !y || !z || !mean || !rstd || M < 0) return RLIPV2_FUSED_EINVAL;
    if (C % 128 != 0 || C > 1024 || C <= 0) return RLIPV2_FUSED_ESHAPE;
    const int grid = (int)(((long long)M + 7) / 8 < kSMs * 8 ? ((long long)M + 7) / 8 : kSMs * 8);
    cudaStream_t s = (cudaStream_t)stream;
#define L(V) case V: add_layernorm_fwd_kernel<V><<<grid, 256, 0, s>>>(x, r, gamma, beta, eps, M, y, z, mean, rstd); break;
    switch (C / 128) { L(1) L(2) L(3) L(4) L(5) L(6) L(7) L(8) }
#undef L
    return done();
}

int rlipv2_layernorm_bwd_acc_f32(const float *dy, const float *z, const float *mean, const float *rstd, const float *gamma,
                                 int M, int C, float *dz, float *dgamma, float *dbeta, int accumulate, void *stream)
{
    if (!dgamma || !dbeta || M < 0) return RLIPV2_FUSED_EINVAL;
    if (C % 128 != 0 || C > 1024 || C <= 0) return RLIPV2_FUSED_ESHAPE;
    cudaStream_t s = (cudaStream_t)stream;
    if (!accumulate) {
        cudaError_t e = cudaMemsetAsync(dgamma, 0, sizeof(float) * C, s);
        if (e == cudaSuccess) e = cudaMemsetAsync(dbeta, 0, sizeof(float) * C, s);
        if (e != cudaSuccess) return (int)e;
    }
    if (M == 0) return 0;
    if (!dy || !z || !mean || !rstd || !gamma || !dz) return RLIPV2_FUSED_EINVAL;
    const int grid = (int)(((long long)M + 7) / 8 < kSMs * 4 ? ((long long)M + 7) / 8 : kSMs * 4);
#define L(V) case V: layernorm_bwd_kernel<V><<<grid, 256, 0, s>>>(dy, z, mean, rstd, gamma, M, dz, dgamma, dbeta); break;
    switch (C / 128) { L(1) L(2) L(3) L(4) L(5) L(6) L(7) L(8) }
#undef L
    return done();
}

int rlipv2_layernorm_bwd_f32(const float *dy, const float *z, const float *mean, const float *rstd, const float *gamma,
                             int M, int C, float *dz, float *dgamma, float *dbeta, void *stream)
{
    return rlipv2_layernorm_bwd_acc_f32(dy, z, mean, rstd, gamma, M, C, dz, dgamma, dbeta, 0, stream);
}

int rlipv2_relu_bwd_colsum_acc_f32(const float *g, const float *y, float *gmasked, float *colsum, int M, int N, int accumulate,
                                   void *stream)
{
    if (!colsum || M < 0 || N <= 0 || N % 32 != 0) return N % 32 != 0 ? RLIPV2_FUSED_ESHAPE : RLIPV2_FUSED_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    if (!accumulate) {
        cudaError_t e = cudaMemsetAsync(colsum, 0, sizeof(float) * N, s);
        if (e != cudaSuccess) return (int)e;
    }
    if (M == 0) return 0;
    if (!g || (y && !gmasked)) return RLIPV2_FUSED_EINVAL;
    const int gx = N / 32;
    int gy = (kSMs * 8 + gx - 1) / gx;
    const int max_gy = (M + 31) / 32;
    if (gy > max_gy) gy = max_gy;
    if (gy < 1) gy = 1;
    if (y) relu_bwd_colsum_kernel<1><<<dim3(gx, gy), 256, 0, s>>>(g, y, gmasked, colsum, M, N);
    else relu_bwd_colsum_kernel<0><<<dim3(gx, gy), 256, 0, s>>>(g, nullptr, nullptr, colsum, M, N);
    return done();
}

int rlipv2_relu_bwd_colsum_f32(const float *g, const float *y, float *gmasked, float *colsum, int M, int N, void *stream)
{
    return rlipv2_relu_bwd_colsum_acc_f32(g, y, gmasked, colsum, M, N, 0, stream);
}

int rlipv2_rowmask_bwd_colsum_acc_f32(const float *g, const unsigned char *rowmask, float *gmasked, float *colsum, int M,
                                      int N, int accumulate, void *stream)
{
    if (!colsum || M < 0 || N <= 0 || N % 32 != 0) return N % 32 != 0 ? RLIPV2_FUSED_ESHAPE : RLIPV2_FUSED_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    if (!accumulate) {
        cudaError_t e = cudaMemsetAsync(colsum, 0, sizeof(float) * N, s);
        if (e != cudaSuccess) return (int)e;
    }
    if (M == 0) return 0;
    if (!g || !rowmask || !gmasked) return RLIPV2_FUSED_EINVAL;
    const int gx = N / 32;
    int gy = (kSMs * 8 + gx - 1) / gx;
    const int max_gy = (M + 31) / 32;
    if (gy > max_gy) gy = max_gy;
    if (gy < 1) gy = 1;
    relu_bwd_colsum_kernel<2><<<dim3(gx, gy), 256, 0, s>>>(g, reinterpret_cast<const float *>(rowmask), gmasked, colsum, M, N);
    return done();
}

int rlipv2_rowmask_bwd_colsum_f32(const float *g, const unsigned char *rowmask, float *gmasked, float *colsum, int M,
                                  int N, void *stream)
{
    return rlipv2_rowmask_bwd_colsum_acc_f32(g, rowmask, gmasked, colsum, M, N, 0, stream);
}

int rlipv2_adamw_scaled_f32(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, long long n, double lr,
                            double beta1, double beta2, double eps, double weight_decay, const float *step,
                            const float *grad_scale, void *stream)
{
    return rlipv2_adamw_dev_f32(param, grad, exp_avg, exp_avg_sq, n, lr, nullptr, beta1, beta2, eps, weight_decay, step,
                                grad_scale, nullptr, stream);
}

int rlipv2_adamw_dev_f32(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, long long n, double lr,
                         const float *lr_dev, double beta1, double beta2, double eps, double weight_decay,
                         const float *step, const float *grad_scale, const unsigned *skip_flag, void *stream)
{
    if (n == 0) return 0;
    if (!param || !grad || !exp_avg || !exp_avg_sq || !step || n < 0) return RLIPV2_FUSED_EINVAL;
    if (((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) return RLIPV2_FUSED_EINVAL;
    const long long n4 = n / 4;
    long long blocks = (n4 + 255) / 256;
    if (blocks > kSMs * 16) blocks = kSMs * 16;
    if (blocks < 1) blocks = 1;
    adamw_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n4, n, (float)lr,
                                                                (float)beta1, (float)beta2, (float)(1.0 - beta1),
                                                                (float)(1.0 - beta2), (float)eps, (float)weight_decay, step,
                                                                grad_scale, lr_dev, skip_flag);
    return done();
}

int rlipv2_adamw_f32(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, long long n, double lr,
                     double beta1, double beta2, double eps, double weight_decay, const float *step, void *stream)
{
    return rlipv2_adamw_scaled_f32(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, nullptr,
                                   stream);
}

int rlipv2_gather_chunks_f32(const long long *table, int n_chunks, float *dst, void *stream)
{
    if (n_chunks == 0) return 0;
    if (!table || !dst || n_chunks < 0) return RLIPV2_FUSED_EINVAL;
    gather_chunks_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(table, dst);
    return done();
}

int rlipv2_wait_host_flag(const unsigned *flag, unsigned *seq, unsigned long long timeout_ns, unsigned *err, void *stream)
{
    if (!flag || !seq || !err) return RLIPV2_FUSED_EINVAL;
    wait_host_flag_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag, seq, timeout_ns, err);
    return done();
}

int rlipv2_stamp_globaltimer(unsigned long long *dst, void *stream)
{
    if (!dst) return RLIPV2_FUSED_EINVAL;
    stamp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(dst);
    return done();
}

int rlipv2_box_refine_f32(const float *delta, const float *ref, float eps, long long n, float *y, void *stream)
{
    if (n == 0) return 0;
    if (!delta || !ref || !y || n < 0) return RLIPV2_FUSED_EINVAL;
    long long blocks = (n + 255) / 256;
    if (blocks > kSMs * 8) blocks = kSMs * 8;
    box_refine_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(delta, ref, eps, n, y);
    return done();
}

int rlipv2_box_pair_loss_f32(const float *src, const float *tgt, int rows, float *l1, float *giou_loss, float *dl1,
                             float *dgiou, void *stream)
{
    if (rows == 0) return 0;
    if (!src || !tgt || !l1 || !giou_loss || !dl1 || !dgiou || rows < 0) return RLIPV2_FUSED_EINVAL;
    if (((uintptr_t)src | (uintptr_t)tgt | (uintptr_t)dl1 | (uintptr_t)dgiou) & 15) return RLIPV2_FUSED_EINVAL;
    box_pair_loss_kernel<<<(rows + 127) / 128, 128, 0, (cudaStream_t)stream>>>(src, tgt, rows, l1, giou_loss, dl1, dgiou);
    return done();
}

static inline unsigned sa_threshold(double p) {
    if (p <= 0.0) return 0u;
    double t = p * 4294967296.0;
    return t >= 4294967295.0 ? 4294967295u : (unsigned)t;
}

int rlipv2_short_attention_fwd_f32(const float *q, const float *k, const float *v, const float *mask, int B, int H, int T,
                                   int D, float scale, double dropout_p, const long long *seed, unsigned salt, float *out,
                                   long long *seed_used, void *stream)
{
    if (B == 0 || H == 0 || T == 0) return 0;
    if (D != kSaD || T > kSaT || T < 0) return RLIPV2_FUSED_ESHAPE;
    if (!q || !k || !v || !out || B < 0 || H < 0 || dropout_p < 0.0 || dropout_p >= 1.0) return RLIPV2_FUSED_EINVAL;
    if (dropout_p > 0.0 && !seed) return RLIPV2_FUSED_EINVAL;
    if (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) & 7) return RLIPV2_FUSED_EINVAL;
    const int blocks = (B * H + kSaWarps - 1) / kSaWarps;
    short_attn_fwd_kernel<<<blocks, kSaWarps * 32, 0, (cudaStream_t)stream>>>(
        q, k, v, mask, B, H, T, scale, sa_threshold(dropout_p), (float)(1.0 / (1.0 - dropout_p)), seed, salt, out, seed_used);
    return done();
}

int rlipv2_short_attention_bwd_f32(const float *q, const float *k, const float *v, const float *mask, const float *grad_out,
                                   int B, int H, int T, int D, float scale, double dropout_p, const long long *seed,
                                   unsigned salt, float *dq, float *dk, float *dv, void *stream)
{
    if (B == 0 || H == 0 || T == 0) return 0;
    if (D != kSaD || T > kSaT || T < 0) return RLIPV2_FUSED_ESHAPE;
    if (!q || !k || !v || !grad_out || !dq || !dk || !dv || B < 0 || H < 0 || dropout_p < 0.0 || dropout_p >= 1.0)
        return RLIPV2_FUSED_EINVAL;
    if (dropout_p > 0.0 && !seed) return RLIPV2_FUSED_EINVAL;
    const int blocks = (B * H + kSaWarps - 1) / kSaWarps;
    short_attn_bwd_kernel<<<blocks, kSaWarps * 32, 0, (cudaStream_t)stream>>>(
        q, k, v, mask, grad_out, B, H, T, scale, sa_threshold(dropout_p), (float)(1.0 / (1.0 - dropout_p)), seed, salt, dq,
        dk, dv);
    return done();
}

int rlipv2_groupnorm_tokens_fwd_f32(const float *x, const float *gamma, const float *beta, float eps, int N, int HW,
                                    int C, int G, double *stats, float *out, long long out_batch_stride, float *mean,
                                    float *rstd, void *stream)
{
    if (N == 0 || HW == 0) return 0;
    if (C != kGnC || G != kGnG) return RLIPV2_FUSED_ESHAPE;
    if (!x || !gamma || !beta || !stats || !out || !mean || !rstd || N < 0 || HW < 0 || (out_batch_stride & 3))
        return RLIPV2_FUSED_EINVAL;
    if (((uintptr_t)x | (uintptr_t)out | (uintptr_t)gamma | (uintptr_t)beta) & 15) return RLIPV2_FUSED_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(stats, 0, (size_t)N * kGnG * 2 * sizeof(double), s);
    if (e != cudaSuccess) return (int)e;
    gn_tok_stats_kernel<<<dim3((HW + kGnRows - 1) / kGnRows, N), 256, 0, s>>>(x, HW, stats);
    long long blocks = ((long long)N * HW * 64 + 255) / 256;
    if (blocks > kSMs * 16) blocks = kSMs * 16;
    gn_tok_apply_kernel<<<(int)blocks, 256, 0, s>>>(x, stats, gamma, beta, eps, N, HW, out, out_batch_stride, mean, rstd);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return done();
}

int rlipv2_groupnorm_tokens_bwd_f32(const float *dy, long long dy_batch_stride, const float *x, const float *mean,
                                    const float *rstd, const float *gamma, int N, int HW, int C, int G, double *sums,
                                    float *dx, float *dgamma, float *dbeta, void *stream)
{
    if (N == 0 || HW == 0) return 0;
    if (C != kGnC || G != kGnG) return RLIPV2_FUSED_ESHAPE;
    if (!dy || !x || !mean || !rstd || !gamma || !sums || !dx || !dgamma || !dbeta || N < 0 || HW < 0 || (dy_batch_stride & 3))
        return RLIPV2_FUSED_EINVAL;
    if (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx | (uintptr_t)gamma) & 15) return RLIPV2_FUSED_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)N * kGnG * 2 * sizeof(double), s);
    if (e != cudaSuccess) return (int)e;
    gn_tok_bwd_stats_kernel<<<dim3((HW + kGnRows - 1) / kGnRows, N), 256, 0, s>>>(dy, dy_batch_stride, x, mean, rstd, gamma,
                                                                                HW, sums, dgamma, dbeta);
    long long blocks = ((long long)N * HW * 64 + 255) / 256;
    if (blocks > kSMs * 16) blocks = kSMs * 16;
    gn_tok_bwd_apply_kernel<<<(int)blocks, 256, 0, s>>>(dy, dy_batch_stride, x, mean, rstd, gamma, sums, N, HW, dx);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return done();
}

int rlipv2_sine_embed_f32(const float *pos, int rows, int n, float *out, void *stream)
{
    if (rows == 0) return 0;
    if (!pos || !out || rows < 0) return RLIPV2_FUSED_EINVAL;
    if (n != 2 && n != 4) return RLIPV2_FUSED_ESHAPE;
    long long blocks = (long long)rows * n;
    if (blocks > kSMs * 16) blocks = kSMs * 16;
    sine_embed_kernel<<<(int)blocks, 128, 0, (cudaStream_t)stream>>>(pos, rows, n, out);
    return done();
}

const char *rlipv2_fused_error_string(int code)
{
    switch (code) {
        case 0: return "success";
        case RLIPV2_FUSED_EINVAL: return "rlipv2_fused: invalid argument";
        case RLIPV2_FUSED_ESHAPE: return "rlipv2_fused: unsupported shape";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "rlipv2_fused: unknown error";
    }
}

unsigned long long rlipv2_fused_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
