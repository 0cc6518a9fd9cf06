// msda_merge.h - merging the grad_value reductions of one (image, query, head) "pair" before they leave the SM.
//
// The fast backward (msda.cu, msda_bwd_d32_l4p4) is bound by the SM -> L2 reduction path: every sampling point scatters
// 4 corner rows of 128 bytes with red.global.add.v4.f32 (reference: one scalar atomicAdd per corner and channel,
// /root/reference/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:125,134,143,152).  Every one of those contributions is
//     grad_value[corner cell, :] += (bilinear weight * attention weight) * grad_out[n, q, m, :]
// i.e. a SCALAR times the pair's one grad_out row.  Two corners of the same pair that fall on the same cell of the same
// level therefore merge by adding their scalars - no vector arithmetic, no exchange of rows.  The four points of one level
// sit one cell apart along the head's direction at initialisation (ms_deform_attn.py:66-76) and stay clustered when
// trained, so a level's 16 corners cover 10-13 distinct cells; corners whose merged scalar is exactly zero (integer sample
// positions: fractional part 0) need no reduction at all.
//
// Ownership rule (deterministic): among the valid corners of a level that fall on one cell, the one with the lowest
// (point, corner) index owns the cell and carries the sum of all their scalars; the others carry 0 and are not issued.
// Coordinates, not addresses, are compared, which makes the test separable in rows and columns:
//     corner (cy, cx) of point q  ==  corner (cy + dh, cx + dw) of point p,   dh = h_q - h_p, dw = w_q - w_p
// and since every scalar factorises as (row weight * attention) * (column weight), so does the sum over p's corners.
//
// The functions are __host__ __device__ so that tests/msda_merge_host_shim.cpp can check them against a brute-force
// cell map on the build box (tests/test_msda_merge_core.py).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MSDA_HD __host__ __device__ __forceinline__
#else
#define MSDA_HD inline
#endif

// One sampling point on its level: unclamped low corner (h, w), validity bits as msda.cu::pack_point
// (bit0 row h inside, bit1 row h+1 inside, bit2 column w inside, bit3 column w+1 inside; 0 when the reference's
// cuh:288 range test fails), fractional parts and attention weight.
struct MsdaPoint {
    int h, w;
    unsigned bits;
    float lh, lw, a;
};

MSDA_HD MsdaPoint msda_point(float loc_w, float loc_h, float a, int H, int W)
{
    // cuh:285-288
    const float h_im = loc_h * (float)H - 0.5f;
    const float w_im = loc_w * (float)W - 0.5f;
    const bool inside = (h_im > -1.f) && (w_im > -1.f) && (h_im < (float)H) && (w_im < (float)W);
    // cuh:40-46
    const float hf = floorf(h_im), wf = floorf(w_im);
    MsdaPoint p;
    p.h = inside ? (int)hf : 0;
    p.w = inside ? (int)wf : 0;
    p.lh = inside ? h_im - hf : 0.f;
    p.lw = inside ? w_im - wf : 0.f;
    p.a = a;
    // cuh:56-78 corner guards
    p.bits = 0;
    if (inside)
        p.bits = (p.h >= 0 ? 1u : 0u) | (p.h + 1 <= H - 1 ? 2u : 0u) | (p.w >= 0 ? 4u : 0u) | (p.w + 1 <= W - 1 ? 8u : 0u);
    return p;
}

// row weights x attention and column weights of a point, zero for rows / columns outside the level
struct MsdaFactors { float r0, r1, c0, c1; };

MSDA_HD MsdaFactors msda_factors(const MsdaPoint &p)
{
    MsdaFactors f;
    f.r0 = (p.bits & 1u) ? (1.f - p.lh) * p.a : 0.f;
    f.r1 = (p.bits & 2u) ? p.lh * p.a : 0.f;
    f.c0 = (p.bits & 4u) ? 1.f - p.lw : 0.f;
    f.c1 = (p.bits & 8u) ? p.lw : 0.f;
    return f;
}

// two-bit validity mask `m` (bit i = line i of the other point is inside) seen from a point whose low line lies d lines
// further: bit j of the result = the other point has a valid line at this point's line j.
MSDA_HD unsigned msda_shift2(unsigned m, int d)
{
    return d == 0 ? m : d == 1 ? (m >> 1) : d == -1 ? ((m << 1) & 3u) : 0u;
}

// Accumulate into tot[cy*2+cx] what point p contributes to the cells of q's corners; when p precedes q, mark the corners
// of q whose cell p owns (dead bit cy*2+cx).
MSDA_HD void msda_merge_from(const MsdaPoint &q, const MsdaPoint &p, const MsdaFactors &fp, bool p_is_lower,
                             float (&tot)[4], unsigned &dead)
{
    const int dh = q.h - p.h, dw = q.w - p.w;
    // p's factor on q's row 0 / row 1 and column 0 / column 1
    const float r0 = dh == 0 ? fp.r0 : dh == 1 ? fp.r1 : 0.f;
    const float r1 = dh == 0 ? fp.r1 : dh == -1 ? fp.r0 : 0.f;
    const float c0 = dw == 0 ? fp.c0 : dw == 1 ? fp.c1 : 0.f;
    const float c1 = dw == 0 ? fp.c1 : dw == -1 ? fp.c0 : 0.f;
    tot[0] = fmaf(r0, c0, tot[0]);
    tot[1] = fmaf(r0, c1, tot[1]);
    tot[2] = fmaf(r1, c0, tot[2]);
    tot[3] = fmaf(r1, c1, tot[3]);
    if (p_is_lower) {
        const unsigned rh = msda_shift2(p.bits & 3u, dh), ch = msda_shift2((p.bits >> 2) & 3u, dw);
        dead |= ((rh & 1u) ? ch : 0u) | ((rh & 2u) ? (ch << 2) : 0u);
    }
}

// Merged reduction scalars of point q (corner order of msda.cu: k = 0 (h, w), 1 (h, w+1), 2 (h+1, w), 3 (h+1, w+1)) given
// the other three points of its level; lower[i] = o[i] precedes q in the level's point order.  A corner that is outside,
// or whose cell is owned by a preceding point, gets exactly 0.f.
MSDA_HD void msda_merge_point(const MsdaPoint &q, const MsdaPoint &o0, bool lower0, const MsdaPoint &o1, bool lower1,
                              const MsdaPoint &o2, bool lower2, float (&s)[4])
{
    const MsdaFactors fq = msda_factors(q);
    float tot[4] = {fq.r0 * fq.c0, fq.r0 * fq.c1, fq.r1 * fq.c0, fq.r1 * fq.c1};
    unsigned dead = 0;
    msda_merge_from(q, o0, msda_factors(o0), lower0, tot, dead);
    msda_merge_from(q, o1, msda_factors(o1), lower1, tot, dead);
    msda_merge_from(q, o2, msda_factors(o2), lower2, tot, dead);
    const unsigned rq = q.bits & 3u, cq = (q.bits >> 2) & 3u;
    const unsigned valid = ((rq & 1u) ? cq : 0u) | ((rq & 2u) ? (cq << 2) : 0u);
    const unsigned live = valid & ~dead;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int k = 0; k < 4; ++k) s[k] = ((live >> k) & 1u) ? tot[k] : 0.f;
}

// What one lane of the backward does: it prepared points (2 odd, 2 odd + 1) of a level (`mine0`, `mine1`), its partner lane
// the other two (`part0`, `part1`); odd = 1 when the partner's points precede this lane's.
MSDA_HD void msda_merge_lane(const MsdaPoint &mine0, const MsdaPoint &mine1, const MsdaPoint &part0, const MsdaPoint &part1,
                             bool odd, float (&s0)[4], float (&s1)[4])
{
    msda_merge_point(mine0, mine1, false, part0, odd, part1, odd, s0);
    msda_merge_point(mine1, mine0, true, part0, odd, part1, odd, s1);
}
