// msda.cu - multi-scale deformable attention (MSDeformAttn) forward / backward for sm_100a.
//
// What it computes (reference: /root/reference/models/ops/src/cuda/ms_deform_im2col_cuda.cuh
// :238-299 forward, :302-403 + :88-159 backward; host wrappers ms_deform_attn_cuda.cu:20-153):
//
//   out[n,q,m,:] = sum_{l<L, p<P} A[n,q,m,l,p] * bilinear(value[n, level l, :, m, :], loc[n,q,m,l,p])
//
// bilinear = zero-padded, align_corners=False sampling at (x*W-0.5, y*H-0.5).
//
// Design (see DESIGN.md "MSDA kernels"):
//  * a "pair" is one (n,q,m): 16 sampling points x 4 corners, each corner one 128-byte row of
//    D=32 fp32 channels.  The op is a gather: per pair 8 KB move through the SM's 128 B/clk
//    L1 data path (64 wavefronts), versus 320 B of streaming input/output.
//  * fast path (fp32, D=32, L=4, P=4 - every ParSeDA call): 8 lanes x float4 own one pair, a
//    warp owns 4 pairs (same head, 4 consecutive queries -> the four groups hit neighbouring
//    cells and share L1 lines).  Phase 1: each lane prepares 2 points of its pair (clamped corner
//    offset + validity bits, fractional parts, attention weight: 4 words) into shared memory.
//    Phase 2: every lane walks the 16 points: 1 LDS.128 + 4 LDG.128 + 16 FFMA per point.
//    Measured (ncu, profiles/msda_r01.md): 90 L1 wavefronts and 231 issue slots per pair versus
//    the reference kernel's 64 + ~560; the kernel runs at 71 % of the L1 data-pipe peak.
//  * a CTA walks a run of 32 consecutive queries head by head, so co-resident warps work on the
//    same head's neighbourhood (L1 reuse) rather than on 8 different heads.
//  * backward fast path: same ownership; d_k = <v_k, grad_out> dot products per corner give
//    grad_attn / grad_loc partials with 16 FFMA instead of ~60; partials are kept in registers
//    for 8 points at a time and reduced across the 8 lanes with a 7-shuffle reduce-scatter per
//    quantity (the reference: 2 __syncthreads + a serial 32-term sum by thread 0 per point);
//    grad_value uses 128-bit vector reductions (red.global.add.v4.f32 -> REDG.E.ADD.F32x4),
//    4x fewer atomic instructions than the reference's scalar atomicAdd.  The backward is bound
//    by the SM->L2 reduction path (1 sector per clock per SM, 256 sector-cycles per pair).
//  * generic path (any D / L / P, fp32 + fp64): plain SIMT, one thread per output element
//    (forward) or one warp per pair (backward); kept for API completeness (gradcheck shapes).
//
// No CPU fallback exists on purpose.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "rlipv2_msda.h"
#include "msda_merge.h"

namespace {

std::atomic<unsigned long long> g_launches{0};
// backward of the fast path: 0 = one reduction per valid corner, 2 CTAs per SM (round 1's kernel), 1 = corners of one pair that
// fall on the same cell merged before they are issued (msda_merge.h), 2 = by call shape: merged for encoder-shaped calls,
// unmerged at 3 CTAs per SM for the plain op's decoder-shaped calls; 3 / 4 = the two 3-CTA variants forced (measurements);
// set through rlipv2_msda_set_backward_mode
std::atomic<int> g_bwd_mode{2};
constexpr int kMergeMinQueries = 8192;

// Measured on B200 (profiles/msda_r02.md, r02u): merging pays where a pair's points cluster - the encoder's self-attention,
// one query per cell sampling around itself (481.8 -> 387.9 us with the initialisation's offsets, 472 -> 460 us with a cell of
// noise on top) - and costs 4 % on decoder-shaped calls (a few hundred queries, learned far-apart points: nothing merges and
// the kernel is DRAM-latency-bound).  The same query-count test selects the forward's variant.
inline bool use_merged_backward(int NQ) {
    const int mode = g_bwd_mode.load(std::memory_order_relaxed);
    return mode == 1 || (mode == 2 && NQ >= kMergeMinQueries);
}

constexpr int kFastD = 32;
constexpr int kFastL = 4;
constexpr int kFastP = 4;
constexpr int kFastLP = kFastL * kFastP;           // 16 points per pair
constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kQueriesPerCta = 32;                  // 8 warps x 4 lane-groups

__device__ __forceinline__ float4 ldg_f4(const float *p) {
    return __ldg(reinterpret_cast<const float4 *>(p));
}

__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Per-level table staged in shared memory by every CTA (the reference re-reads the int64 device
// tensors for every output element, cuh:274-277).
struct LevelRow { int H, W; uint32_t start; uint32_t row_stride; };   // row_stride = W*M*D

// One sampling point as the consumer lanes see it: 4 words = one LDS.128.
//   word 0: element offset of the clamped low corner (a multiple of 32) | validity bits
//           bit0 = row h_low inside, bit1 = row h_low+1 inside, bit2 = col w_low inside,
//           bit3 = col w_low+1 inside (all 0 when the reference's cuh:288 test fails)
//   word 1: lh, word 2: lw (fractional parts), word 3: attention weight
// Corner k offsets follow from the bits: the column step is M*D iff both columns are inside
// (otherwise the two columns coincide after clamping), same for the row step.
__device__ __forceinline__ uint4 pack_point(float loc_w, float loc_h, float a, int H, int W,
                                            uint32_t cell0, uint32_t MD) {
    // cuh:285-288
    const float h_im = loc_h * (float)H - 0.5f;
    const float w_im = loc_w * (float)W - 0.5f;
    const bool inside = (h_im > -1.f) && (w_im > -1.f) && (h_im < (float)H) && (w_im < (float)W);
    // cuh:40-46
    const float hf = floorf(h_im), wf = floorf(w_im);
    const int h_low = inside ? (int)hf : 0, w_low = inside ? (int)wf : 0;
    const float lh = inside ? h_im - hf : 0.f, lw = inside ? w_im - wf : 0.f;
    // cuh:56-78 corner guards
    unsigned bits = 0;
    if (inside) {
        bits = (h_low >= 0 ? 1u : 0u) | (h_low + 1 <= H - 1 ? 2u : 0u) |
               (w_low >= 0 ? 4u : 0u) | (w_low + 1 <= W - 1 ? 8u : 0u);
    }
    const int hl = max(h_low, 0), wl = max(w_low, 0);
    const uint32_t base = cell0 + (uint32_t)(hl * W + wl) * MD;
    return make_uint4(base | bits, __float_as_uint(lh), __float_as_uint(lw), __float_as_uint(a));
}

// ---------------------------------------------------------------------------------------------
// Fast path: fp32, D=32, L=4, P=4.
//
// Schedule: CTA b walks queries [32b, 32b+32) of the flattened (batch, query) axis head by head;
// heads are sliced over grid.y when the query axis alone cannot fill 148 SMs (decoder-shaped
// calls).  A persistent 2-D-tiled schedule for the encoder (Lq == S) was measured in round 1: it
// lifts the L1 hit rate from 68 % to 79 % but not the speed (the kernel is bound by L1 data-pipe
// wavefronts, not by misses) - see DESIGN.md; the code is in git history (commit "MSDA v4").
// ---------------------------------------------------------------------------------------------
struct WarpCtx {
    int grp, sub;             // lane group (pair slot) and lane within the group
    uint32_t MD;              // M * 32
    uint32_t rs0, rs1, rs2, rs3;   // per-level row stride in elements
};

__device__ __forceinline__ void load_level_table(LevelRow *lvl_tab, const int64_t *shapes,
                                                 const int64_t *lsi, uint32_t MD) {
    if (threadIdx.x < kFastL) {
        LevelRow r;
        r.H = (int)shapes[2 * threadIdx.x];
        r.W = (int)shapes[2 * threadIdx.x + 1];
        r.start = (uint32_t)lsi[threadIdx.x];
        r.row_stride = (uint32_t)r.W * MD;
        lvl_tab[threadIdx.x] = r;
    }
    __syncthreads();
}

// Where the 16 (location, weight) of a pair come from.
//   plain op (the reference's native boundary): sampling_loc [NQ, M, L, P, 2] + attn_weight [NQ, M, L, P];
//   fused prologue (PROJ): the raw projection of the query, proj [NQ, M*L*P*3] = M*L*P*2 sampling
//   offsets followed by M*L*P attention logits (one GEMM with the two nn.Linear weights stacked), and
//   the 2-d reference points ref [NQ, L, 2].  The kernel then does what ms_deform_attn.py:102-109
//   does with ~6 elementwise passes over [NQ, M, L, P, 2] tensors: softmax over the pair's 16 logits
//   and loc = ref + offset / (W_l, H_l).
struct PairSrc {
    const float *loc, *attn;
    const float *proj, *ref;
};

__device__ __forceinline__ float group_max8(float x) {
    x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 1));
    x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 2));
    return fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 4));
}

__device__ __forceinline__ float group_sum8(float x) {
    x += __shfl_xor_sync(0xffffffffu, x, 1);
    x += __shfl_xor_sync(0xffffffffu, x, 2);
    return x + __shfl_xor_sync(0xffffffffu, x, 4);
}

// This lane's two points (2*sub, 2*sub+1 - both on level sub/2) of pair (Q, m): locations l4 =
// (x0, y0, x1, y1) and attention weights a2.  Must be called by all 32 lanes (group shuffles).
template <int PROJ>
__device__ __forceinline__ void lane_points(const PairSrc &src, const WarpCtx &c, const LevelRow &my,
                                            int Q, int M, int m, bool live, float4 &l4, float2 &a2)
{
    l4 = make_float4(0.f, 0.f, 0.f, 0.f);
    a2 = make_float2(0.f, 0.f);
    if (!PROJ) {
        if (live) {
            const size_t pair = (size_t)Q * M + m;
            l4 = ldg_f4(src.loc + pair * (kFastLP * 2) + c.sub * 4);
            a2 = __ldg(reinterpret_cast<const float2 *>(src.attn + pair * kFastLP + c.sub * 2));
        }
        return;
    }
    float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float2 lg = make_float2(0.f, 0.f);
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);          // (x, y[, w, h]) of this lane's level
    if (live) {
        const float *row = src.proj + (size_t)Q * (size_t)(M * kFastLP * 3);
        o4 = ldg_f4(row + m * (kFastLP * 2) + c.sub * 4);
        lg = __ldg(reinterpret_cast<const float2 *>(row + M * (kFastLP * 2) + m * kFastLP + c.sub * 2));
        if (PROJ == 2) {
            r = ldg_f4(src.ref + ((size_t)Q * kFastL + (c.sub >> 1)) * 4);
        } else {
            const float2 r2 = __ldg(reinterpret_cast<const float2 *>(src.ref + ((size_t)Q * kFastL + (c.sub >> 1)) * 2));
            r.x = r2.x; r.y = r2.y;
        }
    }
    // softmax over the 16 logits of the pair (F.softmax(.., -1), ms_deform_attn.py:104)
    const float mx = group_max8(fmaxf(lg.x, lg.y));
    const float e0 = expf(lg.x - mx), e1 = expf(lg.y - mx);
    const float sum = group_sum8(e0 + e1);
    a2 = make_float2(__fdiv_rn(e0, sum), __fdiv_rn(e1, sum));
    if (PROJ == 2) {
        // ms_deform_attn.py:110-112: loc = ref[..., :2] + offsets / n_points * ref[..., 2:] * 0.5
        // (/4 and *0.5 are exact; the one rounding of the product and the one of the sum are torch's)
        l4 = make_float4(r.x + __fmul_rn(o4.x * 0.25f, r.z) * 0.5f, r.y + __fmul_rn(o4.y * 0.25f, r.w) * 0.5f,
                         r.x + __fmul_rn(o4.z * 0.25f, r.z) * 0.5f, r.y + __fmul_rn(o4.w * 0.25f, r.w) * 0.5f);
    } else {
        // ms_deform_attn.py:106-109: loc = ref[:, :, None, :, None, :] + offsets / (W_l, H_l)
        const float Wf = (float)my.W, Hf = (float)my.H;
        l4 = make_float4(r.x + __fdiv_rn(o4.x, Wf), r.y + __fdiv_rn(o4.y, Hf),
                         r.x + __fdiv_rn(o4.z, Wf), r.y + __fdiv_rn(o4.w, Hf));
    }
}

// forward for the 4 pairs (same head m, queries Q of the 4 lane groups) owned by this warp
template <int UNROLL, int PROJ>
__device__ __forceinline__ void fwd_warp_pairs(const float *__restrict__ value, const PairSrc &src,
                                               float *__restrict__ out, const WarpCtx &c,
                                               const LevelRow &my, uint4 *mine, int Q, bool live,
                                               int n, int S, int M, int m)
{
    const size_t pair = (size_t)(live ? Q : 0) * M + m;
    // ---- phase 1: this lane prepares points 2*sub, 2*sub+1 (both on level sub/2)
    float4 l4;
    float2 a2;
    lane_points<PROJ>(src, c, my, Q, M, m, live, l4, a2);
    const uint32_t cell0 = ((uint32_t)n * (uint32_t)S + my.start) * c.MD + (uint32_t)m * kFastD;
    mine[c.grp * kFastLP + c.sub * 2 + 0] = pack_point(l4.x, l4.y, a2.x, my.H, my.W, cell0, c.MD);
    mine[c.grp * kFastLP + c.sub * 2 + 1] = pack_point(l4.z, l4.w, a2.y, my.H, my.W, cell0, c.MD);
    __syncwarp();
    // ---- phase 2: gather; lane owns channels sub*4 .. sub*4+3 of its group's pair
    const float *vbase = value + c.sub * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll UNROLL
    for (int p = 0; p < kFastLP; ++p) {
        const uint4 pt = mine[c.grp * kFastLP + p];
        const uint32_t rs = (p < 4) ? c.rs0 : (p < 8) ? c.rs1 : (p < 12) ? c.rs2 : c.rs3;
        const uint32_t bits = pt.x & 15u, base = pt.x & ~31u;
        const uint32_t dx = ((bits & 12u) == 12u) ? c.MD : 0u;
        const uint32_t dy = ((bits & 3u) == 3u) ? rs : 0u;
        const float4 v1 = ldg_f4(vbase + base);
        const float4 v2 = ldg_f4(vbase + base + dx);
        const float4 v3 = ldg_f4(vbase + base + dy);
        const float4 v4 = ldg_f4(vbase + base + dy + dx);
        const float lh = __uint_as_float(pt.y), lw = __uint_as_float(pt.z);
        const float a = __uint_as_float(pt.w);
        // row weights x attention (zero for rows outside), column weights (zero outside)
        const float ha = (bits & 1u) ? (1.f - lh) * a : 0.f;
        const float la = (bits & 2u) ? lh * a : 0.f;
        const float hw = (bits & 4u) ? 1.f - lw : 0.f;
        const float lwv = (bits & 8u) ? lw : 0.f;
        const float w1 = ha * hw, w2 = ha * lwv, w3 = la * hw, w4 = la * lwv;
        acc.x = fmaf(w1, v1.x, acc.x); acc.y = fmaf(w1, v1.y, acc.y);
        acc.z = fmaf(w1, v1.z, acc.z); acc.w = fmaf(w1, v1.w, acc.w);
        acc.x = fmaf(w2, v2.x, acc.x); acc.y = fmaf(w2, v2.y, acc.y);
        acc.z = fmaf(w2, v2.z, acc.z); acc.w = fmaf(w2, v2.w, acc.w);
        acc.x = fmaf(w3, v3.x, acc.x); acc.y = fmaf(w3, v3.y, acc.y);
        acc.z = fmaf(w3, v3.z, acc.z); acc.w = fmaf(w3, v3.w, acc.w);
        acc.x = fmaf(w4, v4.x, acc.x); acc.y = fmaf(w4, v4.y, acc.y);
        acc.z = fmaf(w4, v4.z, acc.z); acc.w = fmaf(w4, v4.w, acc.w);
    }
    if (live)
        *reinterpret_cast<float4 *>(out + pair * kFastD + c.sub * 4) = acc;
    __syncwarp();   // the next call overwrites the staging area
}

// reduce-scatter of x[0..7] over the 8 lanes of a group: lane `sub` returns the group sum of x[sub].
__device__ __forceinline__ float group_reduce_scatter8(const float (&x)[8], int sub) {
    float y[4], z[2];
    const bool b2 = sub & 4, b1 = sub & 2, b0 = sub & 1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float keep = b2 ? x[j + 4] : x[j];
        const float send = b2 ? x[j] : x[j + 4];
        y[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float keep = b1 ? y[j + 2] : y[j];
        const float send = b1 ? y[j] : y[j + 2];
        z[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    const float keep = b0 ? z[1] : z[0], send = b0 ? z[0] : z[1];
    return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

// backward for the 4 pairs owned by this warp.  The 16 points are walked in two halves of 8 so
// that the per-point partials (3 x 8 registers) stay small; scaleW/scaleH = (W, H) of the level
// of point `sub` in each half (cuh:156-158).
// MERGE: phase 1 also forms, per point, the four reduction scalars after merging the corners of the pair that share a cell
// (msda_merge.h: each lane merges its own two points against the level's other two, held by its neighbour lane) and stages
// them as a second 4-word record; phase 2 issues a reduction only for a non-zero scalar.
template <int PROJ, int MERGE>
__device__ __forceinline__ void bwd_warp_pairs(const float *__restrict__ value, const PairSrc &src,
                                               const float *__restrict__ grad_out,
                                               float *__restrict__ grad_value,
                                               float *__restrict__ grad_loc,
                                               float *__restrict__ grad_attn, const WarpCtx &c,
                                               const LevelRow &my, const float (&scaleW)[2],
                                               const float (&scaleH)[2], uint4 *mine, float4 *mine_s,
                                               int Q, bool live, int n, int S, int M, int m)
{
    const size_t pair = (size_t)(live ? Q : 0) * M + m;
    float4 l4;
    float2 a2;
    float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
    lane_points<PROJ>(src, c, my, Q, M, m, live, l4, a2);
    if (live) g4 = ldg_f4(grad_out + pair * kFastD + c.sub * 4);
    const uint32_t cell0 = ((uint32_t)n * (uint32_t)S + my.start) * c.MD + (uint32_t)m * kFastD;
    if (MERGE) {
        MsdaPoint m0 = msda_point(l4.x, l4.y, a2.x, my.H, my.W);
        MsdaPoint m1 = msda_point(l4.z, l4.w, a2.y, my.H, my.W);
        if (!live) { m0.bits = 0u; m1.bits = 0u; }   // dead group: no loads, no reductions
        // the level's other two points live in the neighbour lane (sub ^ 1); w + 1 >= 0 rides with the validity bits
        MsdaPoint q0, q1;
        const unsigned wb0 = ((unsigned)(m0.w + 1) << 4) | m0.bits, wb1 = ((unsigned)(m1.w + 1) << 4) | m1.bits;
        const unsigned xb0 = __shfl_xor_sync(0xffffffffu, wb0, 1), xb1 = __shfl_xor_sync(0xffffffffu, wb1, 1);
        q0.h = __shfl_xor_sync(0xffffffffu, m0.h, 1);   q1.h = __shfl_xor_sync(0xffffffffu, m1.h, 1);
        q0.w = (int)(xb0 >> 4) - 1;                     q1.w = (int)(xb1 >> 4) - 1;
        q0.bits = xb0 & 15u;                            q1.bits = xb1 & 15u;
        q0.lh = __shfl_xor_sync(0xffffffffu, m0.lh, 1); q1.lh = __shfl_xor_sync(0xffffffffu, m1.lh, 1);
        q0.lw = __shfl_xor_sync(0xffffffffu, m0.lw, 1); q1.lw = __shfl_xor_sync(0xffffffffu, m1.lw, 1);
        q0.a = __shfl_xor_sync(0xffffffffu, m0.a, 1);   q1.a = __shfl_xor_sync(0xffffffffu, m1.a, 1);
        float s0[4], s1[4];
        msda_merge_lane(m0, m1, q0, q1, (c.sub & 1) != 0, s0, s1);
        const uint32_t b0 = cell0 + (uint32_t)(max(m0.h, 0) * my.W + max(m0.w, 0)) * c.MD;
        const uint32_t b1 = cell0 + (uint32_t)(max(m1.h, 0) * my.W + max(m1.w, 0)) * c.MD;
        mine[c.grp * kFastLP + c.sub * 2 + 0] =
            make_uint4(b0 | m0.bits, __float_as_uint(m0.lh), __float_as_uint(m0.lw), __float_as_uint(m0.a));
        mine[c.grp * kFastLP + c.sub * 2 + 1] =
            make_uint4(b1 | m1.bits, __float_as_uint(m1.lh), __float_as_uint(m1.lw), __float_as_uint(m1.a));
        mine_s[c.grp * kFastLP + c.sub * 2 + 0] = make_float4(s0[0], s0[1], s0[2], s0[3]);
        mine_s[c.grp * kFastLP + c.sub * 2 + 1] = make_float4(s1[0], s1[1], s1[2], s1[3]);
    } else {
        uint4 p0 = pack_point(l4.x, l4.y, a2.x, my.H, my.W, cell0, c.MD);
        uint4 p1 = pack_point(l4.z, l4.w, a2.y, my.H, my.W, cell0, c.MD);
        if (!live) { p0.x &= ~15u; p1.x &= ~15u; }       // dead group: no loads, no reductions
        mine[c.grp * kFastLP + c.sub * 2 + 0] = p0;
        mine[c.grp * kFastLP + c.sub * 2 + 1] = p1;
    }
    __syncwarp();

    const float *vbase = value + c.sub * 4;
    float *gbase = grad_value + c.sub * 4;
    float ka0 = 0.f, ka1 = 0.f, kx0 = 0.f, kx1 = 0.f, ky0 = 0.f, ky1 = 0.f;   // PROJ: this lane's point of each half
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        float pa[8], pw[8], ph[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint4 pt = mine[c.grp * kFastLP + half * 8 + j];
            const uint32_t rs = half ? ((j < 4) ? c.rs2 : c.rs3) : ((j < 4) ? c.rs0 : c.rs1);
            const uint32_t bits = pt.x & 15u, base = pt.x & ~31u;
            const uint32_t dx = ((bits & 12u) == 12u) ? c.MD : 0u;
            const uint32_t dy = ((bits & 3u) == 3u) ? rs : 0u;
            const bool c1 = (bits & 5u) == 5u, c2 = (bits & 9u) == 9u;
            const bool c3 = (bits & 6u) == 6u, c4 = (bits & 10u) == 10u;
            const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 v1 = c1 ? ldg_f4(vbase + base) : zero;
            const float4 v2 = c2 ? ldg_f4(vbase + base + dx) : zero;
            const float4 v3 = c3 ? ldg_f4(vbase + base + dy) : zero;
            const float4 v4 = c4 ? ldg_f4(vbase + base + dy + dx) : zero;
            const float lh = __uint_as_float(pt.y), lw = __uint_as_float(pt.z);
            const float a = __uint_as_float(pt.w);
            const float hh = 1.f - lh, hw = 1.f - lw;
            // d_k = <v_k, grad_out> over this lane's 4 channels
            const float d1 = fmaf(v1.x, g4.x, fmaf(v1.y, g4.y, fmaf(v1.z, g4.z, v1.w * g4.w)));
            const float d2 = fmaf(v2.x, g4.x, fmaf(v2.y, g4.y, fmaf(v2.z, g4.z, v2.w * g4.w)));
            const float d3 = fmaf(v3.x, g4.x, fmaf(v3.y, g4.y, fmaf(v3.z, g4.z, v3.w * g4.w)));
            const float d4 = fmaf(v4.x, g4.x, fmaf(v4.y, g4.y, fmaf(v4.z, g4.z, v4.w * g4.w)));
            const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
            // cuh:155   grad_attn = top_grad * (w1 v1 + w2 v2 + w3 v3 + w4 v4)
            pa[j] = fmaf(w1, d1, fmaf(w2, d2, fmaf(w3, d3, w4 * d4)));
            // cuh:121-151 grad_w_weight = hh (v2 - v1) + lh (v4 - v3)
            //             grad_h_weight = hw (v3 - v1) + lw (v4 - v2)
            pw[j] = a * fmaf(hh, d2 - d1, lh * (d4 - d3));
            ph[j] = a * fmaf(hw, d3 - d1, lw * (d4 - d2));
            // cuh:125,134,143,152  grad_value[corner] += w_k * top_grad * attn
            if (MERGE) {
                // merged scalars: 0 for a corner that is outside, owned by an earlier corner of the pair, or weightless
                const float4 sw = mine_s[c.grp * kFastLP + half * 8 + j];
                if (sw.x != 0.f) red_add_v4(gbase + base, sw.x * g4.x, sw.x * g4.y, sw.x * g4.z, sw.x * g4.w);
                if (sw.y != 0.f) red_add_v4(gbase + base + dx, sw.y * g4.x, sw.y * g4.y, sw.y * g4.z, sw.y * g4.w);
                if (sw.z != 0.f) red_add_v4(gbase + base + dy, sw.z * g4.x, sw.z * g4.y, sw.z * g4.z, sw.z * g4.w);
                if (sw.w != 0.f) red_add_v4(gbase + base + dy + dx, sw.w * g4.x, sw.w * g4.y, sw.w * g4.z, sw.w * g4.w);
            } else {
                const float s1 = w1 * a, s2 = w2 * a, s3 = w3 * a, s4 = w4 * a;
                if (c1) red_add_v4(gbase + base, s1 * g4.x, s1 * g4.y, s1 * g4.z, s1 * g4.w);
                if (c2) red_add_v4(gbase + base + dx, s2 * g4.x, s2 * g4.y, s2 * g4.z, s2 * g4.w);
                if (c3) red_add_v4(gbase + base + dy, s3 * g4.x, s3 * g4.y, s3 * g4.z, s3 * g4.w);
                if (c4) red_add_v4(gbase + base + dy + dx, s4 * g4.x, s4 * g4.y, s4 * g4.z, s4 * g4.w);
            }
        }
        const float ga = group_reduce_scatter8(pa, c.sub);
        const float gw = group_reduce_scatter8(pw, c.sub);
        const float gh = group_reduce_scatter8(ph, c.sub);
        if (PROJ) {
            float gx, gy;
            if (PROJ == 2) {
                // 4-d reference points: d loc / d offset = (w, h) * 0.5 / n_points; this lane's point half*8 + sub
                // lies on level half*2 + sub/4 (torch: grad_loc * 0.5, * ref_wh, / 4 - one rounding each side)
                float2 wh = make_float2(0.f, 0.f);
                if (live)
                    wh = __ldg(reinterpret_cast<const float2 *>(
                        src.ref + ((size_t)Q * kFastL + (half * 2 + (c.sub >> 2))) * 4 + 2));
                gx = __fmul_rn((gw * scaleW[half]) * 0.5f, wh.x) * 0.25f;
                gy = __fmul_rn((gh * scaleH[half]) * 0.5f, wh.y) * 0.25f;
            } else {
                // d loc / d offset = 1 / (W, H): torch forms grad_loc (cuh:156-158) and divides it again
                gx = __fdiv_rn(gw * scaleW[half], scaleW[half]);
                gy = __fdiv_rn(gh * scaleH[half], scaleH[half]);
            }
            if (half == 0) { ka0 = ga; kx0 = gx; ky0 = gy; } else { ka1 = ga; kx1 = gx; ky1 = gy; }
        } else if (live) {
            const int pt_idx = half * 8 + c.sub;
            *reinterpret_cast<float2 *>(grad_loc + pair * (kFastLP * 2) + pt_idx * 2) =
                make_float2(gw * scaleW[half], gh * scaleH[half]);
            grad_attn[pair * kFastLP + pt_idx] = ga;
        }
    }
    if (PROJ) {
        // softmax backward over the pair's 16 points: dlogit_p = A_p (dA_p - sum_j A_j dA_j); this lane
        // holds points sub and 8 + sub.  grad_loc doubles as grad_proj [NQ, M*L*P*3] here.
        const float A0 = __uint_as_float(mine[c.grp * kFastLP + c.sub].w);
        const float A1 = __uint_as_float(mine[c.grp * kFastLP + 8 + c.sub].w);
        const float dot = group_sum8(fmaf(A0, ka0, A1 * ka1));
        if (live) {
            float *row = grad_loc + (size_t)Q * (size_t)(M * kFastLP * 3);
            float *goff = row + m * (kFastLP * 2);
            float *glog = row + M * (kFastLP * 2) + m * kFastLP;
            *reinterpret_cast<float2 *>(goff + c.sub * 2) = make_float2(kx0, ky0);
            *reinterpret_cast<float2 *>(goff + (8 + c.sub) * 2) = make_float2(kx1, ky1);
            glog[c.sub] = A0 * (ka0 - dot);
            glog[8 + c.sub] = A1 * (ka1 - dot);
        }
    }
    __syncwarp();
}

__device__ __forceinline__ WarpCtx make_ctx(const LevelRow *lvl_tab, int M) {
    WarpCtx c;
    const int lane = threadIdx.x & 31;
    c.grp = lane >> 3;
    c.sub = lane & 7;
    c.MD = (uint32_t)M * kFastD;
    c.rs0 = lvl_tab[0].row_stride; c.rs1 = lvl_tab[1].row_stride;
    c.rs2 = lvl_tab[2].row_stride; c.rs3 = lvl_tab[3].row_stride;
    return c;
}

// ---- linear schedule --------------------------------------------------------------------------
template <int UNROLL, int MINB, int PROJ = 0>
__global__ void __launch_bounds__(kThreads, MINB)
msda_fwd_d32_l4p4(const float *__restrict__ value, const int64_t *__restrict__ shapes,
                  const int64_t *__restrict__ lsi, const float *__restrict__ loc,
                  const float *__restrict__ attn, int NQ, int Lq, int S, int M,
                  float *__restrict__ out)
{
    // PROJ: `loc` carries proj [NQ, M*48] and `attn` carries ref [NQ, 4, 2] (PROJ == 1) or [NQ, 4, 4] (PROJ == 2)
    const PairSrc src = PROJ ? PairSrc{nullptr, nullptr, loc, attn} : PairSrc{loc, attn, nullptr, nullptr};
    __shared__ __align__(16) uint4 prep[kWarpsPerCta][4 * kFastLP];
    __shared__ LevelRow lvl_tab[kFastL];
    load_level_table(lvl_tab, shapes, lsi, (uint32_t)M * kFastD);
    const WarpCtx c = make_ctx(lvl_tab, M);
    const LevelRow my = lvl_tab[c.sub >> 1];
    const int warp = threadIdx.x >> 5;
    const int Q = blockIdx.x * kQueriesPerCta + warp * 4 + c.grp;   // flattened (n,q)
    const bool live = Q < NQ;
    const int n = live ? Q / Lq : 0;
    const int heads_per = (M + gridDim.y - 1) / gridDim.y;
    const int m_begin = blockIdx.y * heads_per;
    const int m_end = min(M, m_begin + heads_per);
    for (int m = m_begin; m < m_end; ++m)
        fwd_warp_pairs<UNROLL, PROJ>(value, src, out, c, my, prep[warp], Q, live, n, S, M, m);
}

template <int MINB, int PROJ = 0, int MERGE = 0>
__global__ void __launch_bounds__(kThreads, MINB)
msda_bwd_d32_l4p4(const float *__restrict__ value, const int64_t *__restrict__ shapes,
                  const int64_t *__restrict__ lsi, const float *__restrict__ loc,
                  const float *__restrict__ attn, const float *__restrict__ grad_out,
                  int NQ, int Lq, int S, int M, float *__restrict__ grad_value,
                  float *__restrict__ grad_loc, float *__restrict__ grad_attn)
{
    // PROJ: `loc` = proj, `attn` = ref, `grad_loc` = grad_proj [NQ, M*48] (fully written), grad_attn unused
    const PairSrc src = PROJ ? PairSrc{nullptr, nullptr, loc, attn} : PairSrc{loc, attn, nullptr, nullptr};
    __shared__ __align__(16) uint4 prep[kWarpsPerCta][4 * kFastLP];
    __shared__ __align__(16) float4 prep_s[MERGE ? kWarpsPerCta : 1][MERGE ? 4 * kFastLP : 1];
    __shared__ LevelRow lvl_tab[kFastL];
    load_level_table(lvl_tab, shapes, lsi, (uint32_t)M * kFastD);
    const WarpCtx c = make_ctx(lvl_tab, M);
    const LevelRow my = lvl_tab[c.sub >> 1];
    const float scaleW[2] = {(float)lvl_tab[c.sub >> 2].W, (float)lvl_tab[2 + (c.sub >> 2)].W};
    const float scaleH[2] = {(float)lvl_tab[c.sub >> 2].H, (float)lvl_tab[2 + (c.sub >> 2)].H};
    const int warp = threadIdx.x >> 5;
    const int Q = blockIdx.x * kQueriesPerCta + warp * 4 + c.grp;
    const bool live = Q < NQ;
    const int n = live ? Q / Lq : 0;
    const int heads_per = (M + gridDim.y - 1) / gridDim.y;
    const int m_begin = blockIdx.y * heads_per;
    const int m_end = min(M, m_begin + heads_per);
    for (int m = m_begin; m < m_end; ++m)
        bwd_warp_pairs<PROJ, MERGE>(value, src, grad_out, grad_value, grad_loc, grad_attn, c, my, scaleW,
                                    scaleH, prep[warp], prep_s[MERGE ? warp : 0], Q, live, n, S, M, m);
}


// ---- experimental: coarse levels staged in shared memory by TMA (north_star's literal design) ------------------------------
// One CTA = (block of queries, head m, image n), 1024 threads.  The cells of levels 2 and 3 of (n, m) - one contiguous
// cell range [coarse_start, S) of `value`, 128 bytes per cell at a 1 KB stride - are fetched by 3-D TMA boxes
// (32 channels x 1 head x 64 cells) into a dense [cells][32] shared-memory tile; the sampling loop then reads the corners of
// levels 2 / 3 from the tile (LDS.128) and those of levels 0 / 1 from global memory as before.  Measured against the default
// kernel in profiles/msda_r02.md: the gather is bound by L1 / shared-memory data-pipe wavefronts, which a corner read costs
// on either path, while the 170 KB tile takes the L1 capacity away from levels 0 / 1 - kept as a documented variant.
constexpr int kTmaThreads = 1024;
constexpr int kTmaBoxCells = 64;

__device__ __forceinline__ uint32_t smem_addr_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__global__ void __launch_bounds__(kTmaThreads, 1)
msda_fwd_tma_d32_l4p4(const __grid_constant__ CUtensorMap tm_v, const float *__restrict__ value,
                      const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
                      const float *__restrict__ loc, const float *__restrict__ attn, int Lq, int S, int M, int qb,
                      int coarse_start, int tile_cells, float *__restrict__ out)
{
    extern __shared__ __align__(128) uint8_t tma_smem[];
    float *tile = reinterpret_cast<float *>(tma_smem);                               // [tile_cells][32]
    uint4 *prep = reinterpret_cast<uint4 *>(tma_smem + (size_t)tile_cells * 128);    // [32 warps][64]
    uint64_t *bar = reinterpret_cast<uint64_t *>(prep + 32 * 4 * kFastLP);
    __shared__ LevelRow lvl_tab[kFastL];
    const int m = blockIdx.y, n = blockIdx.z;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    load_level_table(lvl_tab, shapes, lsi, (uint32_t)M * kFastD);                    // (__syncthreads inside)
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                     :: "r"(smem_addr_u32(bar)), "r"((uint32_t)tile_cells * 128u) : "memory");
        for (int i = 0; i < tile_cells / kTmaBoxCells; ++i)
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                :: "r"(smem_addr_u32(tile + (size_t)i * kTmaBoxCells * 32)), "l"(&tm_v), "r"(smem_addr_u32(bar)), "r"(0), "r"(m),
                   "r"(n * S + coarse_start + i * kTmaBoxCells) : "memory");
    }
    const WarpCtx c = make_ctx(lvl_tab, M);
    const LevelRow my = lvl_tab[c.sub >> 1];
    const int warp = threadIdx.x >> 5;
    uint4 *mine = prep + warp * (4 * kFastLP);
    const bool coarse_lane = (c.sub >> 1) >= 2;                                      // this lane prepares points of level 2 / 3
    {   // wait for the tile
        uint32_t done = 0;
        for (uint32_t spins = 0; !done; ++spins) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(smem_addr_u32(bar)), "r"(0) : "memory");
            if (spins > (1u << 24)) __trap();
        }
    }
    const int q_end = min(Lq, (int)(blockIdx.x + 1) * qb);
    const float *vbase = value + c.sub * 4;
    const float *tbase = tile + c.sub * 4;
    const uint32_t rs2s = (uint32_t)lvl_tab[2].W * kFastD, rs3s = (uint32_t)lvl_tab[3].W * kFastD;
    for (int q0 = blockIdx.x * qb + warp * 4; q0 < q_end; q0 += (kTmaThreads / 32) * 4) {
        const int q = q0 + c.grp;
        const bool live = q < q_end;
        const int Q = n * Lq + (live ? q : q0);
        const size_t pair = (size_t)Q * M + m;
        float4 l4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float2 a2 = make_float2(0.f, 0.f);
        if (live) {
            l4 = ldg_f4(loc + pair * (kFastLP * 2) + c.sub * 4);
            a2 = __ldg(reinterpret_cast<const float2 *>(attn + pair * kFastLP + c.sub * 2));
        }
        // global element offset (levels 0 / 1) or tile element offset (levels 2 / 3) of the level's first cell
        const uint32_t cell0 = coarse_lane ? (my.start - (uint32_t)coarse_start) * kFastD
                                           : ((uint32_t)n * (uint32_t)S + my.start) * c.MD + (uint32_t)m * kFastD;
        const uint32_t md = coarse_lane ? (uint32_t)kFastD : c.MD;
        mine[c.grp * kFastLP + c.sub * 2 + 0] = pack_point(l4.x, l4.y, a2.x, my.H, my.W, cell0, md);
        mine[c.grp * kFastLP + c.sub * 2 + 1] = pack_point(l4.z, l4.w, a2.y, my.H, my.W, cell0, md);
        __syncwarp();
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int p = 0; p < kFastLP; ++p) {
            const uint4 pt = mine[c.grp * kFastLP + p];
            const bool sm = p >= 8;                                                  // compile-time after unrolling
            const uint32_t rs = (p < 4) ? c.rs0 : (p < 8) ? c.rs1 : (p < 12) ? rs2s : rs3s;
            const uint32_t bits = pt.x & 15u, base = pt.x & ~31u;
            const uint32_t dx = ((bits & 12u) == 12u) ? (sm ? (uint32_t)kFastD : c.MD) : 0u;
            const uint32_t dy = ((bits & 3u) == 3u) ? rs : 0u;
            float4 v1, v2, v3, v4;
            if (sm) {
                v1 = *reinterpret_cast<const float4 *>(tbase + base);
                v2 = *reinterpret_cast<const float4 *>(tbase + base + dx);
                v3 = *reinterpret_cast<const float4 *>(tbase + base + dy);
                v4 = *reinterpret_cast<const float4 *>(tbase + base + dy + dx);
            } else {
                v1 = ldg_f4(vbase + base);
                v2 = ldg_f4(vbase + base + dx);
                v3 = ldg_f4(vbase + base + dy);
                v4 = ldg_f4(vbase + base + dy + dx);
            }
            const float lh = __uint_as_float(pt.y), lw = __uint_as_float(pt.z);
            const float a = __uint_as_float(pt.w);
            const float ha = (bits & 1u) ? (1.f - lh) * a : 0.f;
            const float la = (bits & 2u) ? lh * a : 0.f;
            const float hw = (bits & 4u) ? 1.f - lw : 0.f;
            const float lwv = (bits & 8u) ? lw : 0.f;
            const float w1 = ha * hw, w2 = ha * lwv, w3 = la * hw, w4 = la * lwv;
            acc.x = fmaf(w1, v1.x, acc.x); acc.y = fmaf(w1, v1.y, acc.y);
            acc.z = fmaf(w1, v1.z, acc.z); acc.w = fmaf(w1, v1.w, acc.w);
            acc.x = fmaf(w2, v2.x, acc.x); acc.y = fmaf(w2, v2.y, acc.y);
            acc.z = fmaf(w2, v2.z, acc.z); acc.w = fmaf(w2, v2.w, acc.w);
            acc.x = fmaf(w3, v3.x, acc.x); acc.y = fmaf(w3, v3.y, acc.y);
            acc.z = fmaf(w3, v3.z, acc.z); acc.w = fmaf(w3, v3.w, acc.w);
            acc.x = fmaf(w4, v4.x, acc.x); acc.y = fmaf(w4, v4.y, acc.y);
            acc.z = fmaf(w4, v4.z, acc.z); acc.w = fmaf(w4, v4.w, acc.w);
        }
        if (live)
            *reinterpret_cast<float4 *>(out + pair * kFastD + c.sub * 4) = acc;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// Generic path: any channels / levels / points, float or double.
// ---------------------------------------------------------------------------------------------
template <typename T>
struct Geo {
    int64_t o1, o2, o3, o4;
    T lh, lw;
    unsigned mask;
};

template <typename T>
__device__ __forceinline__ Geo<T> generic_geom(T loc_w, T loc_h, int H, int W, int64_t MD) {
    Geo<T> g;
    const T h_im = loc_h * (T)H - (T)0.5;
    const T w_im = loc_w * (T)W - (T)0.5;
    const bool inside = (h_im > (T)-1) && (w_im > (T)-1) && (h_im < (T)H) && (w_im < (T)W);
    const T hf = floor(h_im), wf = floor(w_im);
    int h_low = inside ? (int)hf : 0, w_low = inside ? (int)wf : 0;
    g.lh = inside ? h_im - hf : (T)0;
    g.lw = inside ? w_im - wf : (T)0;
    const bool h0 = h_low >= 0, h1 = h_low + 1 <= H - 1, w0 = w_low >= 0, w1 = w_low + 1 <= W - 1;
    g.mask = inside ? ((h0 && w0) ? 1u : 0u) | ((h0 && w1) ? 2u : 0u) |
                      ((h1 && w0) ? 4u : 0u) | ((h1 && w1) ? 8u : 0u) : 0u;
    const int64_t hl = max(h_low, 0), hh = min(h_low + 1, H - 1);
    const int64_t wl = max(w_low, 0), wh = min(w_low + 1, W - 1);
    g.o1 = (hl * W + wl) * MD; g.o2 = (hl * W + wh) * MD;
    g.o3 = (hh * W + wl) * MD; g.o4 = (hh * W + wh) * MD;
    return g;
}

// forward: one thread per output element (pair, channel); consecutive threads = consecutive channels
template <typename T>
__global__ void __launch_bounds__(256)
msda_fwd_generic(const T *__restrict__ value, const int64_t *__restrict__ shapes,
                 const int64_t *__restrict__ lsi, const T *__restrict__ loc,
                 const T *__restrict__ attn, int64_t total, int Lq, int S, int M, int D, int L,
                 int P, T *__restrict__ out)
{
    const int64_t MD = (int64_t)M * D;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % D);
        const int64_t pair = idx / D;
        const int m = (int)(pair % M);
        const int64_t n = pair / M / Lq;
        const T *lp = loc + pair * L * P * 2;
        const T *ap = attn + pair * L * P;
        T acc = 0;
        for (int l = 0; l < L; ++l) {
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
            const T *v = value + (n * S + lsi[l]) * MD + (int64_t)m * D + c;
            for (int p = 0; p < P; ++p, lp += 2, ++ap) {
                const Geo<T> g = generic_geom<T>(lp[0], lp[1], H, W, MD);
                const T hh = (T)1 - g.lh, hw = (T)1 - g.lw;
                const T v1 = (g.mask & 1u) ? v[g.o1] : (T)0, v2 = (g.mask & 2u) ? v[g.o2] : (T)0;
                const T v3 = (g.mask & 4u) ? v[g.o3] : (T)0, v4 = (g.mask & 8u) ? v[g.o4] : (T)0;
                acc += (hh * hw * v1 + hh * g.lw * v2 + g.lh * hw * v3 + g.lh * g.lw * v4) * ap[0];
            }
        }
        out[idx] = acc;
    }
}

// backward: one warp per pair, lanes stride over channels; warp-shuffle reduction of the
// per-point partials, atomicAdd into grad_value.
template <typename T>
__global__ void __launch_bounds__(256)
msda_bwd_generic(const T *__restrict__ value, const int64_t *__restrict__ shapes,
                 const int64_t *__restrict__ lsi, const T *__restrict__ loc,
                 const T *__restrict__ attn, const T *__restrict__ grad_out, int64_t pairs,
                 int Lq, int S, int M, int D, int L, int P, T *__restrict__ grad_value,
                 T *__restrict__ grad_loc, T *__restrict__ grad_attn)
{
    const int64_t MD = (int64_t)M * D;
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t pair = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; pair < pairs;
         pair += warps) {
        const int m = (int)(pair % M);
        const int64_t n = pair / M / Lq;
        const T *lp = loc + pair * L * P * 2;
        const T *ap = attn + pair * L * P;
        T *glp = grad_loc + pair * L * P * 2;
        T *gap = grad_attn + pair * L * P;
        const T *go = grad_out + pair * D;
        for (int l = 0; l < L; ++l) {
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
            const int64_t base = (n * S + lsi[l]) * MD + (int64_t)m * D;
            for (int p = 0; p < P; ++p, lp += 2, ++ap, glp += 2, ++gap) {
                const Geo<T> g = generic_geom<T>(lp[0], lp[1], H, W, MD);
                const T a = ap[0];
                const T hh = (T)1 - g.lh, hw = (T)1 - g.lw;
                const T w1 = hh * hw, w2 = hh * g.lw, w3 = g.lh * hw, w4 = g.lh * g.lw;
                T pa = 0, pw = 0, ph = 0;
                for (int c = lane; c < D; c += 32) {
                    const T tg = go[c];
                    const T *v = value + base + c;
                    T *gv = grad_value + base + c;
                    const T v1 = (g.mask & 1u) ? v[g.o1] : (T)0, v2 = (g.mask & 2u) ? v[g.o2] : (T)0;
                    const T v3 = (g.mask & 4u) ? v[g.o3] : (T)0, v4 = (g.mask & 8u) ? v[g.o4] : (T)0;
                    const T tga = tg * a;
                    pa += tg * (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);
                    pw += tga * (hh * (v2 - v1) + g.lh * (v4 - v3));
                    ph += tga * (hw * (v3 - v1) + g.lw * (v4 - v2));
                    if (g.mask & 1u) atomicAdd(gv + g.o1, w1 * tga);
                    if (g.mask & 2u) atomicAdd(gv + g.o2, w2 * tga);
                    if (g.mask & 4u) atomicAdd(gv + g.o3, w3 * tga);
                    if (g.mask & 8u) atomicAdd(gv + g.o4, w4 * tga);
                }
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) {
                    pa += __shfl_xor_sync(0xffffffffu, pa, s);
                    pw += __shfl_xor_sync(0xffffffffu, pw, s);
                    ph += __shfl_xor_sync(0xffffffffu, ph, s);
                }
                if (lane == 0) {
                    gap[0] = pa;
                    glp[0] = pw * (T)W;
                    glp[1] = ph * (T)H;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
inline bool bad_dims(int batch, int spatial_size, int num_heads, int channels, int num_levels,
                     int num_query, int num_point) {
    return batch < 0 || spatial_size < 0 || num_query < 0 || num_heads <= 0 || channels <= 0 ||
           num_levels <= 0 || num_point <= 0;
}

inline bool fast_ok(int batch, int spatial_size, int num_heads, int channels, int num_levels,
                    int num_query, int num_point) {
    if (channels != kFastD || num_levels != kFastL || num_point != kFastP) return false;
    // 32-bit element offsets into value, 32-bit flattened query index
    const long long velems = (long long)batch * spatial_size * num_heads * channels;
    const long long nq = (long long)batch * num_query;
    return velems < (1ll << 32) && nq < (1ll << 31) - kQueriesPerCta;
}

// the fast kernels use 128-bit ld.global / red.global.add.v4 on value, sampling_loc, out and grad_value and 64-bit loads
// on attn / ref: a contiguous tensor whose storage offset is not 16-byte aligned (a slice / view - the reference's
// scalar kernel accepts it) must take the generic kernel instead of faulting with a misaligned address
inline bool aligned16(const void *a, const void *b = nullptr, const void *c = nullptr, const void *d = nullptr,
                      const void *e = nullptr, const void *f = nullptr, const void *g = nullptr) {
    return ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)c) | ((uintptr_t)d) | ((uintptr_t)e) | ((uintptr_t)f) |
             ((uintptr_t)g)) & 15) == 0;
}

// CTAs along x walk 32 consecutive queries; heads are sliced over grid.y only when the query
// dimension alone leaves SMs idle (decoder-shaped calls: a few hundred queries).
inline dim3 fast_grid(int NQ, int num_heads) {
    const int gx = (NQ + kQueriesPerCta - 1) / kQueriesPerCta;
    int gy = 1;
    while (gx * gy < 2 * 148 && gy < num_heads && num_heads % (gy * 2) == 0) gy *= 2;
    return dim3((unsigned)gx, (unsigned)gy, 1);
}

inline int grid_for(long long work, int per_block) {
    long long b = (work + per_block - 1) / per_block;
    const long long cap = 148ll * 64;     // grid-stride beyond this
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

template <typename T>
int forward_impl(const T *value, const int64_t *shapes, const int64_t *lsi, const T *loc,
                 const T *attn, int batch, int spatial_size, int num_heads, int channels,
                 int num_levels, int num_query, int num_point, T *out, cudaStream_t stream,
                 bool allow_fast)
{
    if (bad_dims(batch, spatial_size, num_heads, channels, num_levels, num_query, num_point))
        return RLIPV2_MSDA_EINVAL;
    const long long pairs = (long long)batch * num_query * num_heads;
    if (pairs == 0) return 0;
    if (!value || !shapes || !lsi || !loc || !attn || !out) return RLIPV2_MSDA_EINVAL;
    if (pairs * channels >= (1ll << 62)) return RLIPV2_MSDA_ETOOBIG;
    if (allow_fast) {
        const int NQ = batch * num_query;
        const dim3 grid = fast_grid(NQ, num_heads);
        // encoder-shaped calls are bound by the L1 data pipe: 5 CTAs/SM of 48-register threads;
        // decoder-shaped calls (random cells, DRAM-bound) want the deeper unroll (16 points of
        // loads in flight).  Both measured on B200, profiles/msda_r01.md.
        if (NQ >= 8192)
            msda_fwd_d32_l4p4<8, 5><<<grid, kThreads, 0, stream>>>(
                (const float *)value, shapes, lsi, (const float *)loc, (const float *)attn, NQ,
                num_query, spatial_size, num_heads, (float *)out);
        else
            msda_fwd_d32_l4p4<16, 3><<<grid, kThreads, 0, stream>>>(
                (const float *)value, shapes, lsi, (const float *)loc, (const float *)attn, NQ,
                num_query, spatial_size, num_heads, (float *)out);
    } else {
        const long long total = pairs * channels;
        msda_fwd_generic<T><<<grid_for(total, 256), 256, 0, stream>>>(
            value, shapes, lsi, loc, attn, total, num_query, spatial_size, num_heads, channels,
            num_levels, num_point, out);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

template <typename T>
int backward_impl(const T *value, const int64_t *shapes, const int64_t *lsi, const T *loc,
                  const T *attn, const T *grad_out, int batch, int spatial_size, int num_heads,
                  int channels, int num_levels, int num_query, int num_point, T *grad_value,
                  T *grad_loc, T *grad_attn, cudaStream_t stream, bool allow_fast)
{
    if (bad_dims(batch, spatial_size, num_heads, channels, num_levels, num_query, num_point))
        return RLIPV2_MSDA_EINVAL;
    const long long velems = (long long)batch * spatial_size * num_heads * channels;
    const long long pairs = (long long)batch * num_query * num_heads;
    if (velems > 0) {
        if (!grad_value) return RLIPV2_MSDA_EINVAL;
        cudaError_t e = cudaMemsetAsync(grad_value, 0, (size_t)velems * sizeof(T), stream);
        if (e != cudaSuccess) return (int)e;
    }
    if (pairs == 0) return 0;
    if (!value || !shapes || !lsi || !loc || !attn || !grad_out || !grad_loc || !grad_attn)
        return RLIPV2_MSDA_EINVAL;
    if (allow_fast) {
        const int NQ = batch * num_query;
        const dim3 grid = fast_grid(NQ, num_heads);
        const int mode = g_bwd_mode.load(std::memory_order_relaxed);
        const float *v = (const float *)value, *l = (const float *)loc, *a = (const float *)attn, *g = (const float *)grad_out;
        float *gv = (float *)grad_value, *gl = (float *)grad_loc, *ga = (float *)grad_attn;
        if (mode == 3)                                    // experiment: merged at 80 registers, 3 CTAs per SM
            msda_bwd_d32_l4p4<3, 0, 1><<<grid, kThreads, 0, stream>>>(v, shapes, lsi, l, a, g, NQ, num_query, spatial_size,
                                                                       num_heads, gv, gl, ga);
        else if (use_merged_backward(NQ))
            msda_bwd_d32_l4p4<2, 0, 1><<<grid, kThreads, 0, stream>>>(v, shapes, lsi, l, a, g, NQ, num_query, spatial_size,
                                                                       num_heads, gv, gl, ga);
        else if (mode == 4 || (mode == 2 && NQ < kMergeMinQueries))
            // decoder-shaped calls scatter over the whole value map and wait on DRAM: 3 CTAs per SM (80 registers, the same
            // source) hide more of that latency - config 5 backward 142.5 -> 129.7 us (r02w); on the encoder call the
            // lower register budget costs 10 % (474 -> 518 us), so mode 2 uses it below kMergeMinQueries only
            msda_bwd_d32_l4p4<3, 0, 0><<<grid, kThreads, 0, stream>>>(v, shapes, lsi, l, a, g, NQ, num_query, spatial_size,
                                                                       num_heads, gv, gl, ga);
        else
            msda_bwd_d32_l4p4<2><<<grid, kThreads, 0, stream>>>(v, shapes, lsi, l, a, g, NQ, num_query, spatial_size,
                                                                 num_heads, gv, gl, ga);
    } else {
        msda_bwd_generic<T><<<grid_for(pairs * 32, 256), 256, 0, stream>>>(
            value, shapes, lsi, loc, attn, grad_out, pairs, num_query, spatial_size, num_heads,
            channels, num_levels, num_point, grad_value, grad_loc, grad_attn);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

// fused-prologue variant: fast path only (fp32, D=32, L=4, P=4, 2-d reference points)
int proj_forward_impl(const float *value, const int64_t *shapes, const int64_t *lsi, const float *ref,
                      const float *proj, int batch, int spatial_size, int num_heads, int channels,
                      int num_levels, int num_query, int num_point, float *out, cudaStream_t stream, int ref_dim = 2)
{
    if (ref_dim != 2 && ref_dim != 4) return RLIPV2_MSDA_EINVAL;
    if (bad_dims(batch, spatial_size, num_heads, channels, num_levels, num_query, num_point))
        return RLIPV2_MSDA_EINVAL;
    if (!fast_ok(batch, spatial_size, num_heads, channels, num_levels, num_query, num_point))
        return RLIPV2_MSDA_ESHAPE;
    const int NQ = batch * num_query;
    if (NQ == 0) return 0;
    if (!value || !shapes || !lsi || !ref || !proj || !out) return RLIPV2_MSDA_EINVAL;
    if (!aligned16(value, ref, proj, out)) return RLIPV2_MSDA_EALIGN;
    const dim3 grid = fast_grid(NQ, num_heads);
    if (ref_dim == 4) {
        if (NQ >= 8192)
            msda_fwd_d32_l4p4<8, 5, 2><<<grid, kThreads, 0, stream>>>(value, shapes, lsi, proj, ref, NQ, num_query,
                                                                     spatial_size, num_heads, out);
        else
            msda_fwd_d32_l4p4<16, 3, 2><<<grid, kThreads, 0, stream>>>(value, shapes, lsi, proj, ref, NQ, num_query,
                                                                      spatial_size, num_heads, out);
    } else if (NQ >= 8192)
        msda_fwd_d32_l4p4<8, 5, 1><<<grid, kThreads, 0, stream>>>(value, shapes, lsi, proj, ref, NQ, num_query,
                                                                 spatial_size, num_heads, out);
    else
        msda_fwd_d32_l4p4<16, 3, 1><<<grid, kThreads, 0, stream>>>(value, shapes, lsi, proj, ref, NQ, num_query,
                                                                  spatial_size, num_heads, out);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

int proj_backward_impl(const float *value, const int64_t *shapes, const int64_t *lsi, const float *ref,
                       const float *proj, const float *grad_out, int batch, int spatial_size, int num_heads,
                       int channels, int num_levels, int num_query, int num_point, float *grad_value,
                       float *grad_proj, cudaStream_t stream, int ref_dim = 2)
{
    if (ref_dim != 2 && ref_dim != 4) return RLIPV2_MSDA_EINVAL;
    if (bad_dims(batch, spatial_size, num_heads, channels, num_levels, num_query, num_point))
        return RLIPV2_MSDA_EINVAL;
    if (!fast_ok(batch, spatial_size, num_heads, channels, num_levels, num_query, num_point))
        return RLIPV2_MSDA_ESHAPE;
    const long long velems = (long long)batch * spatial_size * num_heads * channels;
    if (velems > 0) {
        if (!grad_value) return RLIPV2_MSDA_EINVAL;
        cudaError_t e = cudaMemsetAsync(grad_value, 0, (size_t)velems * sizeof(float), stream);
        if (e != cudaSuccess) return (int)e;
    }
    const int NQ = batch * num_query;
    if (NQ == 0) return 0;
    if (!value || !shapes || !lsi || !ref || !proj || !grad_out || !grad_proj) return RLIPV2_MSDA_EINVAL;
    if (!aligned16(value, ref, proj, grad_out, grad_value, grad_proj)) return RLIPV2_MSDA_EALIGN;
    const dim3 grid = fast_grid(NQ, num_heads);
    const bool merge = use_merged_backward(NQ);
    if (ref_dim == 4) {
        if (merge)
            msda_bwd_d32_l4p4<2, 2, 1><<<grid, kThreads, 0, stream>>>(value, shapes, lsi, proj, ref, grad_out, NQ, num_query,
                                                                       spatial_size, num_heads, grad_value, grad_proj, nullptr);
        else
            msda_bwd_d32_l4p4<2, 2><<<grid, kThreads, 0, stream>>>(value, shapes, lsi, proj, ref, grad_out, NQ, num_query,
                                                                    spatial_size, num_heads, grad_value, grad_proj, nullptr);
    } else if (merge)
        msda_bwd_d32_l4p4<2, 1, 1><<<grid, kThreads, 0, stream>>>(value, shapes, lsi, proj, ref, grad_out, NQ, num_query,
                                                                   spatial_size, num_heads, grad_value, grad_proj, nullptr);
    else
        msda_bwd_d32_l4p4<2, 1><<<grid, kThreads, 0, stream>>>(value, shapes, lsi, proj, ref, grad_out, NQ, num_query,
                                                                spatial_size, num_heads, grad_value, grad_proj, nullptr);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

}  // namespace

extern "C" {

int rlipv2_msda_proj_forward_f32(const float *value, const int64_t *spatial_shapes,
                                 const int64_t *level_start_index, const float *reference_points,
                                 const float *proj, int batch, int spatial_size, int num_heads, int channels,
                                 int num_levels, int num_query, int num_point, float *out, void *stream)
{
    return proj_forward_impl(value, spatial_shapes, level_start_index, reference_points, proj, batch,
                             spatial_size, num_heads, channels, num_levels, num_query, num_point, out,
                             (cudaStream_t)stream);
}

int rlipv2_msda_proj_backward_f32(const float *value, const int64_t *spatial_shapes,
                                  const int64_t *level_start_index, const float *reference_points,
                                  const float *proj, const float *grad_out, int batch, int spatial_size,
                                  int num_heads, int channels, int num_levels, int num_query, int num_point,
                                  float *grad_value, float *grad_proj, void *stream)
{
    return proj_backward_impl(value, spatial_shapes, level_start_index, reference_points, proj, grad_out,
                              batch, spatial_size, num_heads, channels, num_levels, num_query, num_point,
                              grad_value, grad_proj, (cudaStream_t)stream);
}

int rlipv2_msda_proj_ref4_forward_f32(const float *value, const int64_t *spatial_shapes,
                                      const int64_t *level_start_index, const float *reference_boxes,
                                      const float *proj, int batch, int spatial_size, int num_heads, int channels,
                                      int num_levels, int num_query, int num_point, float *out, void *stream)
{
    return proj_forward_impl(value, spatial_shapes, level_start_index, reference_boxes, proj, batch,
                             spatial_size, num_heads, channels, num_levels, num_query, num_point, out,
                             (cudaStream_t)stream, 4);
}

int rlipv2_msda_proj_ref4_backward_f32(const float *value, const int64_t *spatial_shapes,
                                       const int64_t *level_start_index, const float *reference_boxes,
                                       const float *proj, const float *grad_out, int batch, int spatial_size,
                                       int num_heads, int channels, int num_levels, int num_query, int num_point,
                                       float *grad_value, float *grad_proj, void *stream)
{
    return proj_backward_impl(value, spatial_shapes, level_start_index, reference_boxes, proj, grad_out,
                              batch, spatial_size, num_heads, channels, num_levels, num_query, num_point,
                              grad_value, grad_proj, (cudaStream_t)stream, 4);
}

int rlipv2_msda_forward_tma_f32(const float *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                                const float *sampling_loc, const float *attn_weight, int batch, int spatial_size,
                                int num_heads, int channels, int num_levels, int num_query, int num_point,
                                int coarse_start, float *out, void *stream)
{
    if (bad_dims(batch, spatial_size, num_heads, channels, num_levels, num_query, num_point)) return RLIPV2_MSDA_EINVAL;
    if (!fast_ok(batch, spatial_size, num_heads, channels, num_levels, num_query, num_point)) return RLIPV2_MSDA_ESHAPE;
    if (!value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !out) return RLIPV2_MSDA_EINVAL;
    if (!aligned16(value, sampling_loc, attn_weight, out)) return RLIPV2_MSDA_EALIGN;
    if (coarse_start <= 0 || coarse_start >= spatial_size) return RLIPV2_MSDA_EINVAL;
    const int tile_cells = (spatial_size - coarse_start + kTmaBoxCells - 1) / kTmaBoxCells * kTmaBoxCells;
    const size_t smem = (size_t)tile_cells * 128 + 32 * 4 * kFastLP * sizeof(uint4) + 16;
    if (smem > 220 * 1024) return RLIPV2_MSDA_ESHAPE;          // coarse levels too large for one SM's shared memory
    typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                      const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeTiledFn enc = nullptr;
    if (!enc) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return RLIPV2_MSDA_EINVAL;
        enc = reinterpret_cast<EncodeTiledFn>(p);
    }
    // value [N*S cells][M heads][32 channels]: box = 64 cells x 1 head x 32 channels, dense rows of 128 B in shared memory
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)kFastD, (cuuint64_t)num_heads, (cuuint64_t)batch * spatial_size};
    cuuint64_t strides[2] = {(cuuint64_t)kFastD * 4, (cuuint64_t)num_heads * kFastD * 4};
    cuuint32_t box[3] = {(cuuint32_t)kFastD, 1, (cuuint32_t)kTmaBoxCells};
    cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(value), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return RLIPV2_MSDA_EINVAL;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(msda_fwd_tma_d32_l4p4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = smem;
    }
    // four waves of 148 CTAs over (query blocks, heads, images)
    int blocks = (4 * 148) / (num_heads * batch);
    if (blocks < 1) blocks = 1;
    int qb = (num_query + blocks - 1) / blocks;
    qb = (qb + 3) / 4 * 4;
    blocks = (num_query + qb - 1) / qb;
    dim3 grid((unsigned)blocks, (unsigned)num_heads, (unsigned)batch);
    msda_fwd_tma_d32_l4p4<<<grid, kTmaThreads, smem, (cudaStream_t)stream>>>(tm, value, spatial_shapes, level_start_index,
                                                                             sampling_loc, attn_weight, num_query, spatial_size,
                                                                             num_heads, qb, coarse_start, tile_cells, out);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

int rlipv2_msda_forward_f32(const float *value, const int64_t *spatial_shapes,
                            const int64_t *level_start_index, const float *sampling_loc,
                            const float *attn_weight, int batch, int spatial_size, int num_heads,
                            int channels, int num_levels, int num_query, int num_point,
                            float *out, void *stream)
{
    return forward_impl<float>(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                               batch, spatial_size, num_heads, channels, num_levels, num_query,
                               num_point, out, (cudaStream_t)stream,
                               fast_ok(batch, spatial_size, num_heads, channels, num_levels,
                                       num_query, num_point) &&
                                   aligned16(value, sampling_loc, attn_weight, out));
}

int rlipv2_msda_forward_f64(const double *value, const int64_t *spatial_shapes,
                            const int64_t *level_start_index, const double *sampling_loc,
                            const double *attn_weight, int batch, int spatial_size, int num_heads,
                            int channels, int num_levels, int num_query, int num_point,
                            double *out, void *stream)
{
    return forward_impl<double>(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                batch, spatial_size, num_heads, channels, num_levels, num_query,
                                num_point, out, (cudaStream_t)stream, false);
}

int rlipv2_msda_backward_f32(const float *value, const int64_t *spatial_shapes,
                             const int64_t *level_start_index, const float *sampling_loc,
                             const float *attn_weight, const float *grad_out, int batch,
                             int spatial_size, int num_heads, int channels, int num_levels,
                             int num_query, int num_point, float *grad_value,
                             float *grad_sampling_loc, float *grad_attn_weight, void *stream)
{
    return backward_impl<float>(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                grad_out, batch, spatial_size, num_heads, channels, num_levels,
                                num_query, num_point, grad_value, grad_sampling_loc,
                                grad_attn_weight, (cudaStream_t)stream,
                                fast_ok(batch, spatial_size, num_heads, channels, num_levels,
                                        num_query, num_point) &&
                                    aligned16(value, sampling_loc, attn_weight, grad_out, grad_value,
                                              grad_sampling_loc, grad_attn_weight));
}

int rlipv2_msda_backward_f64(const double *value, const int64_t *spatial_shapes,
                             const int64_t *level_start_index, const double *sampling_loc,
                             const double *attn_weight, const double *grad_out, int batch,
                             int spatial_size, int num_heads, int channels, int num_levels,
                             int num_query, int num_point, double *grad_value,
                             double *grad_sampling_loc, double *grad_attn_weight, void *stream)
{
    return backward_impl<double>(value, spatial_shapes, level_start_index, sampling_loc,
                                 attn_weight, grad_out, batch, spatial_size, num_heads, channels,
                                 num_levels, num_query, num_point, grad_value, grad_sampling_loc,
                                 grad_attn_weight, (cudaStream_t)stream, false);
}

const char *rlipv2_msda_error_string(int code)
{
    if (code == 0) return "success";
    if (code == RLIPV2_MSDA_EINVAL) return "rlipv2_msda: invalid argument (dimension or null pointer)";
    if (code == RLIPV2_MSDA_ETOOBIG) return "rlipv2_msda: problem too large";
    if (code == RLIPV2_MSDA_ESHAPE) return "rlipv2_msda: fused-prologue entry points need fp32, D=32, L=4, P=4";
    if (code == RLIPV2_MSDA_EALIGN) return "rlipv2_msda: fused-prologue entry points need 16-byte aligned pointers";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "rlipv2_msda: unknown error";
}

int rlipv2_msda_abi_version(void) { return RLIPV2_MSDA_ABI_VERSION; }

int rlipv2_msda_set_backward_mode(int mode)
{
    if (mode < 0 || mode > 4) return RLIPV2_MSDA_EINVAL;
    g_bwd_mode.store(mode, std::memory_order_relaxed);
    return 0;
}

int rlipv2_msda_get_backward_mode(void) { return g_bwd_mode.load(std::memory_order_relaxed); }


unsigned long long rlipv2_msda_launch_count(void)
{
    return g_launches.load(std::memory_order_relaxed);
}

}  // extern "C"
