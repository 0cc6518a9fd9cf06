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
//    L1 data path, versus 320 B of streaming input/output.
//  * fast path (fp32, D=32, L=4, P=4 - every ParSeDA call): 8 lanes x float4 own one pair, a
//    warp owns 4 pairs (same head, 4 consecutive queries -> the four groups hit neighbouring
//    cells and share L1 lines).  Phase 1: each lane prepares 2 points of its pair (bilinear
//    weights x attention weight, clamped corner offsets) into shared memory.  Phase 2: every
//    lane walks the 16 points: 2 LDS.128 + 4 LDG.128 + 16 FFMA per point.  That keeps the issue
//    slots per pair (~140) below the 64-wavefront L1 floor (256 slots), i.e. the kernel runs at
//    the gather floor instead of the reference's ~560 issue slots per pair.
//  * a CTA walks a run of 32 consecutive queries head by head, so co-resident warps work on the
//    same head's neighbourhood (L1 reuse) rather than on 8 different heads.
//  * backward fast path: same ownership; d_k = <v_k, grad_out> dot products per corner give
//    grad_attn / grad_loc partials with 16 FFMA instead of ~60; partials are kept in registers
//    for all 16 points and reduced across the 8 lanes with a 14-shuffle reduce-scatter per
//    quantity (the reference: 2 __syncthreads + a serial 32-term sum by thread 0 per point);
//    grad_value uses 128-bit vector reductions (red.global.add.v4.f32 -> REDG.E.ADD.F32x4),
//    4x fewer atomic instructions than the reference's scalar atomicAdd.
//  * generic path (any D / L / P, fp32 + fp64): plain SIMT, one thread per output element
//    (forward) or one warp per pair (backward); kept for API completeness (gradcheck shapes).
//
// No CPU fallback exists on purpose.
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "rlipv2_msda.h"

namespace {

std::atomic<unsigned long long> g_launches{0};

constexpr int kFastD = 32;
constexpr int kFastL = 4;
constexpr int kFastP = 4;
constexpr int kFastLP = kFastL * kFastP;           // 16 points per pair
constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kQueriesPerCta = 32;                  // 8 warps x 4 lane-groups
// per warp: 4 pairs x 16 points x (int4 + float4)
constexpr int kPrepBytesPerWarp = 4 * kFastLP * 32;

__device__ __forceinline__ float4 ldg_f4(const float *p) {
    return __ldg(reinterpret_cast<const float4 *>(p));
}

__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// One sampling point, prepared by one lane.
struct PointGeom {
    uint32_t o1, o2, o3, o4;   // element offsets of the 4 corners (clamped into the map)
    float lh, lw;              // fractional parts
    unsigned mask;             // bit k set <=> corner k+1 is inside the map
    bool inside;               // reference's cuh:288 test
};

// Geometry of one sampling point.  Mirrors cuh:285-288 (h_im/w_im and the range test) and
// cuh:40-78 (floor, fractional weights, per-corner guards).  `cell0` = element offset of
// value[n, level_start, m, 0]; `MD` = num_heads*channels.
__device__ __forceinline__ PointGeom point_geom(float loc_w, float loc_h, int H, int W,
                                                uint32_t cell0, uint32_t MD) {
    PointGeom g;
    const float h_im = loc_h * (float)H - 0.5f;
    const float w_im = loc_w * (float)W - 0.5f;
    g.inside = (h_im > -1.f) && (w_im > -1.f) && (h_im < (float)H) && (w_im < (float)W);
    const float hf = floorf(h_im), wf = floorf(w_im);
    int h_low = (int)hf, w_low = (int)wf;
    g.lh = h_im - hf;
    g.lw = w_im - wf;
    if (!g.inside) { h_low = 0; w_low = 0; g.lh = 0.f; g.lw = 0.f; }
    const bool h0 = h_low >= 0, h1 = h_low + 1 <= H - 1;
    const bool w0 = w_low >= 0, w1 = w_low + 1 <= W - 1;
    g.mask = g.inside ? ((h0 && w0) ? 1u : 0u) | ((h0 && w1) ? 2u : 0u) |
                        ((h1 && w0) ? 4u : 0u) | ((h1 && w1) ? 8u : 0u) : 0u;
    const int hl = max(h_low, 0), hh = min(h_low + 1, H - 1);
    const int wl = max(w_low, 0), wh = min(w_low + 1, W - 1);
    const uint32_t r0 = (uint32_t)(hl * W), r1 = (uint32_t)(hh * W);
    g.o1 = cell0 + (r0 + (uint32_t)wl) * MD;
    g.o2 = cell0 + (r0 + (uint32_t)wh) * MD;
    g.o3 = cell0 + (r1 + (uint32_t)wl) * MD;
    g.o4 = cell0 + (r1 + (uint32_t)wh) * MD;
    return g;
}

// ---------------------------------------------------------------------------------------------
// Fast forward: fp32, D=32, L=4, P=4.
// grid.x = ceil(NQ / 32) where NQ = batch*num_query (queries flattened over the batch).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
msda_fwd_d32_l4p4(const float *__restrict__ value, const int64_t *__restrict__ shapes,
                  const int64_t *__restrict__ lsi, const float *__restrict__ loc,
                  const float *__restrict__ attn, int NQ, int Lq, int S, int M,
                  float *__restrict__ out)
{
    __shared__ __align__(16) unsigned char prep_smem[kWarpsPerCta * kPrepBytesPerWarp];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane >> 3, sub = lane & 7;
    int4 *p_off = reinterpret_cast<int4 *>(prep_smem + warp * kPrepBytesPerWarp);
    float4 *p_wt = reinterpret_cast<float4 *>(p_off + 4 * kFastLP);

    // this lane prepares points 2*sub, 2*sub+1 -> both on level sub/2
    const int lvl = sub >> 1;
    const int H = (int)shapes[2 * lvl], W = (int)shapes[2 * lvl + 1];
    const uint32_t lstart = (uint32_t)lsi[lvl];
    const uint32_t MD = (uint32_t)M * kFastD;

    const int Q = blockIdx.x * kQueriesPerCta + warp * 4 + grp;   // flattened (n,q)
    const bool live = Q < NQ;
    const int n = live ? Q / Lq : 0;

    for (int m = 0; m < M; ++m) {
        const size_t pair = (size_t)(live ? Q : 0) * M + m;
        // ---- phase 1: geometry + weights of 2 points per lane -> shared memory
        float4 l4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float2 a2 = make_float2(0.f, 0.f);
        if (live) {
            l4 = ldg_f4(loc + pair * (kFastLP * 2) + sub * 4);
            a2 = __ldg(reinterpret_cast<const float2 *>(attn + pair * kFastLP + sub * 2));
        }
        const uint32_t cell0 = ((uint32_t)n * (uint32_t)S + lstart) * MD + (uint32_t)m * kFastD;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float lw_ = k ? l4.z : l4.x, lh_ = k ? l4.w : l4.y;
            const float a = k ? a2.y : a2.x;
            const PointGeom g = point_geom(lw_, lh_, H, W, cell0, MD);
            const float hh = 1.f - g.lh, hw = 1.f - g.lw;
            float4 wt;
            wt.x = (g.mask & 1u) ? hh * hw * a : 0.f;
            wt.y = (g.mask & 2u) ? hh * g.lw * a : 0.f;
            wt.z = (g.mask & 4u) ? g.lh * hw * a : 0.f;
            wt.w = (g.mask & 8u) ? g.lh * g.lw * a : 0.f;
            const int slot = grp * kFastLP + sub * 2 + k;
            p_off[slot] = make_int4((int)g.o1, (int)g.o2, (int)g.o3, (int)g.o4);
            p_wt[slot] = wt;
        }
        __syncwarp();
        // ---- phase 2: gather.  lane owns channels sub*4 .. sub*4+3 of its group's pair
        const float *vbase = value + sub * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int p = 0; p < kFastLP; ++p) {
            const int4 o = p_off[grp * kFastLP + p];
            const float4 w = p_wt[grp * kFastLP + p];
            const float4 v1 = ldg_f4(vbase + (uint32_t)o.x);
            const float4 v2 = ldg_f4(vbase + (uint32_t)o.y);
            const float4 v3 = ldg_f4(vbase + (uint32_t)o.z);
            const float4 v4 = ldg_f4(vbase + (uint32_t)o.w);
            acc.x = fmaf(w.x, v1.x, acc.x); acc.y = fmaf(w.x, v1.y, acc.y);
            acc.z = fmaf(w.x, v1.z, acc.z); acc.w = fmaf(w.x, v1.w, acc.w);
            acc.x = fmaf(w.y, v2.x, acc.x); acc.y = fmaf(w.y, v2.y, acc.y);
            acc.z = fmaf(w.y, v2.z, acc.z); acc.w = fmaf(w.y, v2.w, acc.w);
            acc.x = fmaf(w.z, v3.x, acc.x); acc.y = fmaf(w.z, v3.y, acc.y);
            acc.z = fmaf(w.z, v3.z, acc.z); acc.w = fmaf(w.z, v3.w, acc.w);
            acc.x = fmaf(w.w, v4.x, acc.x); acc.y = fmaf(w.w, v4.y, acc.y);
            acc.z = fmaf(w.w, v4.z, acc.z); acc.w = fmaf(w.w, v4.w, acc.w);
        }
        if (live)
            *reinterpret_cast<float4 *>(out + pair * kFastD + sub * 4) = acc;
        __syncwarp();   // phase-1 of the next head overwrites the staging area
    }
}

// ---------------------------------------------------------------------------------------------
// Fast backward: fp32, D=32, L=4, P=4.  Same ownership as the forward kernel.
// ---------------------------------------------------------------------------------------------
// reduce-scatter of x[0..15] over the 8 lanes of a group: on return lane `sub` holds the group
// sums of x[2*sub] and x[2*sub+1] in r0, r1.
__device__ __forceinline__ void group_reduce_scatter16(const float (&x)[kFastLP], int sub,
                                                       float &r0, float &r1) {
    float y[8], z[4];
    const bool b2 = sub & 4, b1 = sub & 2, b0 = sub & 1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float keep = b2 ? x[j + 8] : x[j];
        const float send = b2 ? x[j] : x[j + 8];
        y[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float keep = b1 ? y[j + 4] : y[j];
        const float send = b1 ? y[j] : y[j + 4];
        z[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    {
        const float keep0 = b0 ? z[2] : z[0], send0 = b0 ? z[0] : z[2];
        const float keep1 = b0 ? z[3] : z[1], send1 = b0 ? z[1] : z[3];
        r0 = keep0 + __shfl_xor_sync(0xffffffffu, send0, 1);
        r1 = keep1 + __shfl_xor_sync(0xffffffffu, send1, 1);
    }
}

__global__ void __launch_bounds__(kThreads)
msda_bwd_d32_l4p4(const float *__restrict__ value, const int64_t *__restrict__ shapes,
                  const int64_t *__restrict__ lsi, const float *__restrict__ loc,
                  const float *__restrict__ attn, const float *__restrict__ grad_out,
                  int NQ, int Lq, int S, int M, float *__restrict__ grad_value,
                  float *__restrict__ grad_loc, float *__restrict__ grad_attn)
{
    __shared__ __align__(16) unsigned char prep_smem[kWarpsPerCta * kPrepBytesPerWarp];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane >> 3, sub = lane & 7;
    int4 *p_off = reinterpret_cast<int4 *>(prep_smem + warp * kPrepBytesPerWarp);
    float4 *p_geo = reinterpret_cast<float4 *>(p_off + 4 * kFastLP);   // (lh, lw, attn, mask)

    const int lvl = sub >> 1;
    const int H = (int)shapes[2 * lvl], W = (int)shapes[2 * lvl + 1];
    const uint32_t lstart = (uint32_t)lsi[lvl];
    const uint32_t MD = (uint32_t)M * kFastD;

    const int Q = blockIdx.x * kQueriesPerCta + warp * 4 + grp;
    const bool live = Q < NQ;
    const int n = live ? Q / Lq : 0;

    for (int m = 0; m < M; ++m) {
        const size_t pair = (size_t)(live ? Q : 0) * M + m;
        float4 l4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float2 a2 = make_float2(0.f, 0.f);
        float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) {
            l4 = ldg_f4(loc + pair * (kFastLP * 2) + sub * 4);
            a2 = __ldg(reinterpret_cast<const float2 *>(attn + pair * kFastLP + sub * 2));
            g4 = ldg_f4(grad_out + pair * kFastD + sub * 4);
        }
        const uint32_t cell0 = ((uint32_t)n * (uint32_t)S + lstart) * MD + (uint32_t)m * kFastD;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float lw_ = k ? l4.z : l4.x, lh_ = k ? l4.w : l4.y;
            const float a = k ? a2.y : a2.x;
            const PointGeom g = point_geom(lw_, lh_, H, W, cell0, MD);
            const int slot = grp * kFastLP + sub * 2 + k;
            p_off[slot] = make_int4((int)g.o1, (int)g.o2, (int)g.o3, (int)g.o4);
            p_geo[slot] = make_float4(g.lh, g.lw, live ? a : 0.f, __uint_as_float(live ? g.mask : 0u));
        }
        __syncwarp();

        const float *vbase = value + sub * 4;
        float *gbase = grad_value + sub * 4;
        float pa[kFastLP], pw[kFastLP], ph[kFastLP];
#pragma unroll
        for (int p = 0; p < kFastLP; ++p) {
            const int4 o = p_off[grp * kFastLP + p];
            const float4 ge = p_geo[grp * kFastLP + p];
            const unsigned mask = __float_as_uint(ge.w);
            const float lh = ge.x, lw = ge.y, a = ge.z;
            const float hh = 1.f - lh, hw = 1.f - lw;
            const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 v1 = (mask & 1u) ? ldg_f4(vbase + (uint32_t)o.x) : zero;
            const float4 v2 = (mask & 2u) ? ldg_f4(vbase + (uint32_t)o.y) : zero;
            const float4 v3 = (mask & 4u) ? ldg_f4(vbase + (uint32_t)o.z) : zero;
            const float4 v4 = (mask & 8u) ? ldg_f4(vbase + (uint32_t)o.w) : zero;
            // d_k = <v_k, grad_out> over this lane's 4 channels
            const float d1 = fmaf(v1.x, g4.x, fmaf(v1.y, g4.y, fmaf(v1.z, g4.z, v1.w * g4.w)));
            const float d2 = fmaf(v2.x, g4.x, fmaf(v2.y, g4.y, fmaf(v2.z, g4.z, v2.w * g4.w)));
            const float d3 = fmaf(v3.x, g4.x, fmaf(v3.y, g4.y, fmaf(v3.z, g4.z, v3.w * g4.w)));
            const float d4 = fmaf(v4.x, g4.x, fmaf(v4.y, g4.y, fmaf(v4.z, g4.z, v4.w * g4.w)));
            const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
            // cuh:155   grad_attn = top_grad * (w1 v1 + w2 v2 + w3 v3 + w4 v4)
            pa[p] = fmaf(w1, d1, fmaf(w2, d2, fmaf(w3, d3, w4 * d4)));
            // cuh:121-151 grad_w_weight = hh (v2 - v1) + lh (v4 - v3); grad_h_weight = hw (v3 - v1) + lw (v4 - v2)
            pw[p] = a * fmaf(hh, d2 - d1, lh * (d4 - d3));
            ph[p] = a * fmaf(hw, d3 - d1, lw * (d4 - d2));
            // cuh:125,134,143,152  grad_value[corner] += w_k * top_grad * attn
            const float s1 = w1 * a, s2 = w2 * a, s3 = w3 * a, s4 = w4 * a;
            if (mask & 1u) red_add_v4(gbase + (uint32_t)o.x, s1 * g4.x, s1 * g4.y, s1 * g4.z, s1 * g4.w);
            if (mask & 2u) red_add_v4(gbase + (uint32_t)o.y, s2 * g4.x, s2 * g4.y, s2 * g4.z, s2 * g4.w);
            if (mask & 4u) red_add_v4(gbase + (uint32_t)o.z, s3 * g4.x, s3 * g4.y, s3 * g4.z, s3 * g4.w);
            if (mask & 8u) red_add_v4(gbase + (uint32_t)o.w, s4 * g4.x, s4 * g4.y, s4 * g4.z, s4 * g4.w);
        }
        float ga0, ga1, gw0, gw1, gh0, gh1;
        group_reduce_scatter16(pa, sub, ga0, ga1);
        group_reduce_scatter16(pw, sub, gw0, gw1);
        group_reduce_scatter16(ph, sub, gh0, gh1);
        if (live) {
            // cuh:156-158: d/d(loc_w) carries the factor width, d/d(loc_h) the factor height
            *reinterpret_cast<float4 *>(grad_loc + pair * (kFastLP * 2) + sub * 4) =
                make_float4(gw0 * (float)W, gh0 * (float)H, gw1 * (float)W, gh1 * (float)H);
            *reinterpret_cast<float2 *>(grad_attn + pair * kFastLP + sub * 2) = make_float2(ga0, ga1);
        }
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
        const int grid = (NQ + kQueriesPerCta - 1) / kQueriesPerCta;
        msda_fwd_d32_l4p4<<<grid, kThreads, 0, stream>>>(
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
        const int grid = (NQ + kQueriesPerCta - 1) / kQueriesPerCta;
        msda_bwd_d32_l4p4<<<grid, kThreads, 0, stream>>>(
            (const float *)value, shapes, lsi, (const float *)loc, (const float *)attn,
            (const float *)grad_out, NQ, num_query, spatial_size, num_heads, (float *)grad_value,
            (float *)grad_loc, (float *)grad_attn);
    } else {
        msda_bwd_generic<T><<<grid_for(pairs * 32, 256), 256, 0, stream>>>(
            value, shapes, lsi, loc, attn, grad_out, pairs, num_query, spatial_size, num_heads,
            channels, num_levels, num_point, grad_value, grad_loc, grad_attn);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

}  // namespace

extern "C" {

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
                                       num_query, num_point));
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
                                        num_query, num_point));
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
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "rlipv2_msda: unknown error";
}

int rlipv2_msda_abi_version(void) { return RLIPV2_MSDA_ABI_VERSION; }

unsigned long long rlipv2_msda_launch_count(void)
{
    return g_launches.load(std::memory_order_relaxed);
}

}  // extern "C"
