// TEST INFRASTRUCTURE - host build of rlipv2_b200/csrc/msda_merge.h (g++ -shared), the header msda.cu compiles for sm_100a:
// the merging of a pair's grad_value reductions can then be checked on the build box, which has no GPU.
//  * msda_merge_level: the two lanes that own a level's four points, run one after the other exactly as the kernel's
//    phase 1 does (each lane: its own two points against the neighbour lane's two);
//  * msda_merge_grad_value: the grad_value scatter of the whole fast backward (fp32, D = 32, L = 4, P = 4) with the kernel's
//    addressing (clamped low corner + column / row steps derived from the validity bits), sequential, for comparison with
//    the oracle's grad_value.  Not linked into any product library.
#include <stdint.h>

#include "msda_merge.h"

extern "C" {

// loc [4][2] (x, y in [0, 1] units), attn [4] -> s [4][4] merged scalars; geo [4][3] = (h, w, bits) per point
void msda_merge_level(const float *loc, const float *attn, int H, int W, float *s, int *geo)
{
    MsdaPoint p[4];
    for (int i = 0; i < 4; ++i) {
        p[i] = msda_point(loc[2 * i], loc[2 * i + 1], attn[i], H, W);
        geo[3 * i] = p[i].h; geo[3 * i + 1] = p[i].w; geo[3 * i + 2] = (int)p[i].bits;
    }
    float s0[4], s1[4];
    msda_merge_lane(p[0], p[1], p[2], p[3], false, s0, s1);          // even lane: points 0, 1
    for (int k = 0; k < 4; ++k) { s[k] = s0[k]; s[4 + k] = s1[k]; }
    msda_merge_lane(p[2], p[3], p[0], p[1], true, s0, s1);           // odd lane: points 2, 3
    for (int k = 0; k < 4; ++k) { s[8 + k] = s0[k]; s[12 + k] = s1[k]; }
}

// grad_value [N, S, M, 32] (zero-initialised by the caller, double accumulation so that the comparison sees the merge and
// not the summation order); returns the number of reductions issued.
long long msda_merge_grad_value(const int64_t *shapes, const int64_t *lsi, const float *loc, const float *attn,
                                const float *grad_out, int N, int S, int M, int Lq, double *grad_value)
{
    const uint32_t MD = (uint32_t)M * 32u;
    long long issued = 0;
    for (int n = 0; n < N; ++n)
        for (int q = 0; q < Lq; ++q)
            for (int m = 0; m < M; ++m) {
                const size_t pair = ((size_t)n * Lq + q) * M + m;
                const float *g = grad_out + pair * 32;
                for (int l = 0; l < 4; ++l) {
                    const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
                    const uint32_t cell0 = ((uint32_t)n * (uint32_t)S + (uint32_t)lsi[l]) * MD + (uint32_t)m * 32u;
                    const uint32_t rs = (uint32_t)W * MD;
                    MsdaPoint p[4];
                    for (int i = 0; i < 4; ++i)
                        p[i] = msda_point(loc[(pair * 16 + l * 4 + i) * 2], loc[(pair * 16 + l * 4 + i) * 2 + 1],
                                          attn[pair * 16 + l * 4 + i], H, W);
                    float s[4][4];
                    msda_merge_lane(p[0], p[1], p[2], p[3], false, s[0], s[1]);
                    msda_merge_lane(p[2], p[3], p[0], p[1], true, s[2], s[3]);
                    for (int i = 0; i < 4; ++i) {
                        const int hl = p[i].h > 0 ? p[i].h : 0, wl = p[i].w > 0 ? p[i].w : 0;
                        const uint32_t base = cell0 + (uint32_t)(hl * W + wl) * MD;
                        const uint32_t dx = ((p[i].bits & 12u) == 12u) ? MD : 0u;
                        const uint32_t dy = ((p[i].bits & 3u) == 3u) ? rs : 0u;
                        const uint32_t off[4] = {base, base + dx, base + dy, base + dy + dx};
                        for (int k = 0; k < 4; ++k)
                            if (s[i][k] != 0.f) {
                                ++issued;
                                for (int c = 0; c < 32; ++c) grad_value[off[k] + c] += (double)(s[i][k] * g[c]);
                            }
                    }
                }
            }
    return issued;
}

}  // extern "C"
