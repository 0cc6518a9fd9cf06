// attn_tf32.cu - fused softmax-attention cores on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), sm_100a.
//
// Reference arithmetic (all three are bmm -> softmax -> dropout -> bmm chains on fp32 activations):
//   ALIF bidirectional attention   /root/reference/models/fuse_helper.py:395-445      8 heads x 256, Tv 273 x Tl 256
//   RobertaLayer self-attention    models/modeling_roberta.py:185-241                 12 heads x 64, Tl 256, additive key mask
//   decoder query self-attention   models/dab_deformable/deformable_transformer.py:1383-1390   8 heads x 32, 300 / 150 queries
// Interface and constraints: include/rlipv2_attn.h.
//
// attn_fwd_kernel   (6 warps, one CTA per (128-query tile, batch x head, output-column slice), 512 TMEM columns)
//   warp 0 / lane 0  TMA producer: per 32-wide head-dim chunk a [128 x 32] box of Q and the [Nk_pad x 32] rows of K
//                    (128B-swizzled, K-major); then per 32-key chunk the [32 x slice] rows of V as stored (MN-major boxes,
//                    SWIZZLE_128B_ATOM_32B).  One mbarrier ring serves both phases, so V is prefetched under the softmax.
//   warp 1           TMEM allocation; lane 0 issues S = Q K^T (tcgen05.mma kind::tf32, M 128, N <= 256 per instruction)
//                    into TMEM columns [0, Nk_pad), waits for the softmax warps, then O = P V with the A operand read
//                    from those same TMEM columns (tcgen05.mma with a tensor-memory A operand) into columns [Nk_pad, ..)
//   warps 2-5        softmax: thread = query row = TMEM lane; pass 1 row maximum, pass 2 exp / row sum / dropout and the
//                    unnormalised probabilities written back over S (tcgen05.st); epilogue: O x 1 / (sum (1 - p)),
//                    transposed through shared memory into whole 128-byte row segments of out
// attn_bwd_ds_kernel  (one CTA per (128-query tile, batch x head, 128-key chunk), 256 TMEM columns): S tile and
//                    dP = dO V^T tile on the tensor cores, delta = rowsum(dO o O) beside them, then
//                    dS = scale * P o (dP~ - delta) and the dropped probabilities P~ to the [B*H, Tq, Nk_pad] workspaces
// attn_bgemm_kernel   batched C (+)= A B^T with the operands read as stored (K-major or MN-major 3-D TMA boxes):
//                    dQ = dS K, dK = dS^T Q, dV = P~^T dO
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <atomic>

#include "rlipv2_attn.h"

namespace {

std::atomic<unsigned long long> g_launches{0};

constexpr int kThreads = 192;
constexpr int kTile = 128;                  // query rows per CTA = TMEM lanes
constexpr int kChunk = 32;                  // 32 fp32 = 128 bytes = one swizzle row
constexpr int kTileBytes = kTile * kChunk * 4;          // 16 KB: a [128 x 32] fp32 box

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    // a barrier that never flips (byte-count / descriptor bug) must become an error, not a hung GPU
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        :: "r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (see csrc/dense_tf32.cu)
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// MN-major operand tile (stored [contraction rows, MN columns]), SWIZZLE_128B_BASE32B: chunks of [32 rows x 32 columns]
// 4096 B apart (LBO), 4-row swizzle atoms 512 B apart (SBO) (see csrc/dense_tf32.cu, section 6a of DESIGN.md)
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand from tensor memory (lanes = rows, 32-bit columns = K)
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" :: "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
           "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
           "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
           "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void red_add_v4(float *addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) << 4, a / b format TF32 (2) << 7 / << 10,
// bit 15 / 16 = A / B is MN-major, n_dim = N >> 3 at bit 17, m_dim = M >> 4 at bit 24
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);
}

// Dropout: one splitmix64 hash of (seed, salt, group index) decides FOUR consecutive keys of a row (16 bits each against a
// 16-bit threshold), so the mask costs ~6 integer instructions per element instead of ~25.  Group index =
// (batch x head x query row) * (Nk_pad / 4) + key / 4: the forward and the backward regenerate the same bits.
__device__ __forceinline__ unsigned long long hash4(unsigned long long seed, unsigned salt, unsigned idx) {
    unsigned long long z = seed * 0x9E3779B97F4A7C15ull + ((unsigned long long)salt << 32 | idx) + 0xD1B54A32D192ED03ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ bool keep_of(unsigned long long z, int u, unsigned thresh16) {
    return ((unsigned)(z >> (16 * u)) & 0xFFFFu) >= thresh16;
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
constexpr int kSoftmaxWarps = 8;                          // 2 per TMEM lane quadrant: each takes half of the columns
constexpr int kThreadsAttn = 64 + 32 * kSoftmaxWarps;     // + TMA warp + MMA warp

struct FwdArgs {
    int B, H, Tq, Nk, D, dv_tile, Nk_pad, nbox, box_rows, stages, stage_bytes;
    float scale, inv_keep;
    unsigned thresh, salt;
    const float *key_bias;
    float *out;
    long long o_ld, o_bs;
    float *stats;
    const long long *seed;
    long long *seed_used;
};

__global__ void __launch_bounds__(kThreadsAttn, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const FwdArgs a)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + a.stages * a.stage_bytes);
    uint64_t *empty_bar = full_bar + a.stages;
    uint64_t *s_full = empty_bar + a.stages, *p_ready = s_full + 1, *o_full = s_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(s_full + 3);
    float *red = reinterpret_cast<float *>(tmem_slot + 4);               // [2][128]: per-half row max, then row sum
    float *kb = red + 2 * kTile;                                          // [Nk_pad]: log2e * key bias, -inf past Nk

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, bh = blockIdx.y;
    const int b = bh / a.H, h = bh - b * a.H;
    const int col_base = h * a.D;
    const int dv0 = blockIdx.z * a.dv_tile;
    const int nd = a.D / kChunk, nkc = a.Nk_pad / kChunk;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tm_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tm_k) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tm_v) : "memory");
        for (int s = 0; s < a.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(s_full, 1);
        mbar_init(p_ready, 32 * kSoftmaxWarps);
        mbar_init(o_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && a.seed_used) *a.seed_used = a.seed ? *a.seed : 0ll;
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= 2) {
        for (int c = threadIdx.x - 64; c < a.Nk_pad; c += 32 * kSoftmaxWarps)
            kb[c] = c < a.Nk ? (a.key_bias ? kLog2e * __ldg(a.key_bias + (size_t)b * a.Nk + c) : 0.f) : -INFINITY;
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_s = *tmem_slot;                                  // S / P: columns [0, Nk_pad)
    const uint32_t tmem_o = tmem_s + (uint32_t)a.Nk_pad;                 // O: columns [Nk_pad, Nk_pad + dv_tile)

    if (warp == 0) {
        if (lane == 0) {                                                 // ---- TMA producer
            int it = 0;
            for (int dc = 0; dc < nd; ++dc, ++it) {
                const int s = it % a.stages;
                mbar_wait(&empty_bar[s], ((it / a.stages) & 1) ^ 1);
                mbar_expect_tx(&full_bar[s], kTileBytes + a.Nk_pad * 128);
                uint8_t *st = smem + s * a.stage_bytes;
                tma_load_3d(st, &tm_q, &full_bar[s], col_base + dc * kChunk, qt * kTile, b);
                for (int bx = 0; bx < a.nbox; ++bx)
                    tma_load_3d(st + kTileBytes + bx * a.box_rows * 128, &tm_k, &full_bar[s], col_base + dc * kChunk,
                                bx * a.box_rows, b);
            }
            for (int kc = 0; kc < nkc; ++kc, ++it) {
                const int s = it % a.stages;
                mbar_wait(&empty_bar[s], ((it / a.stages) & 1) ^ 1);
                mbar_expect_tx(&full_bar[s], a.dv_tile * 128);
                uint8_t *st = smem + s * a.stage_bytes;
                for (int j = 0; j < a.dv_tile / 32; ++j)
                    tma_load_3d(st + j * 4096, &tm_v, &full_bar[s], col_base + dv0 + j * 32, kc * kChunk, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                                 // ---- MMA issuer
            const uint32_t idesc_s = make_idesc(a.box_rows, false, false);
            const uint32_t idesc_o = make_idesc(a.dv_tile, false, true);
            int it = 0;
            for (int dc = 0; dc < nd; ++dc, ++it) {
                const int s = it % a.stages;
                mbar_wait(&full_bar[s], (it / a.stages) & 1);
                fence_after();
                const uint32_t sa = smem_u32(smem + s * a.stage_bytes);
                const uint32_t sb = sa + kTileBytes;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    for (int bx = 0; bx < a.nbox; ++bx)
                        umma_ss(tmem_s + (uint32_t)(bx * a.box_rows), desc_k_sw128(sa + k * 32),
                                desc_k_sw128(sb + bx * a.box_rows * 128 + k * 32), idesc_s, (dc | k) != 0 ? 1u : 0u);
                umma_commit(&empty_bar[s]);
            }
            umma_commit(s_full);                                         // scores complete
            mbar_wait(p_ready, 0);                                       // probabilities written back to TMEM
            fence_after();
            for (int kc = 0; kc < nkc; ++kc, ++it) {
                const int s = it % a.stages;
                mbar_wait(&full_bar[s], (it / a.stages) & 1);
                fence_after();
                const uint32_t sv = smem_u32(smem + s * a.stage_bytes);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_ts(tmem_o, tmem_s + (uint32_t)(kc * kChunk + k * 8), desc_mn_sw128(sv + k * 1024), idesc_o,
                            (kc | k) != 0 ? 1u : 0u);
                umma_commit(&empty_bar[s]);
            }
            umma_commit(o_full);
        }
    } else {                                                             // ---- softmax + epilogue warps 2..9
        const int q4 = warp & 3;                                         // TMEM lane quadrant this warp may touch
        const int half = (warp - 2) >> 2;                                // which half of the key / output columns
        const int rloc = q4 * 32 + lane;
        const int grow = qt * kTile + rloc;                              // query row of this thread
        const uint32_t lane_s = tmem_s + ((uint32_t)(q4 * 32) << 16);
        const unsigned long long seed = (a.thresh && a.seed) ? (unsigned long long)*a.seed : 0ull;
        const float scale2 = a.scale * kLog2e;
        const int c_lo = half ? (nkc + 1) / 2 : 0, c_hi = half ? nkc : (nkc + 1) / 2;
        mbar_wait(s_full, 0);
        fence_after();
        uint32_t r[32];
        float m = -INFINITY;
#pragma unroll 1
        for (int c = c_lo; c < c_hi; ++c) {
            tmem_ld32(lane_s + (uint32_t)(c * 32), r);
#pragma unroll
            for (int j = 0; j < 32; ++j) m = fmaxf(m, fmaf(__uint_as_float(r[j]), scale2, kb[c * 32 + j]));
        }
        red[half * kTile + rloc] = m;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        m = fmaxf(red[rloc], red[kTile + rloc]);                         // finite: the first half always holds a real key
        asm volatile("bar.sync 1, 256;" ::: "memory");                   // both halves have read before `red` is reused
        float sum = 0.f;
        const unsigned grp_row = (unsigned)((size_t)bh * a.Tq + grow) * (unsigned)(a.Nk_pad / 4);
#pragma unroll 1
        for (int c = c_lo; c < c_hi; ++c) {
            tmem_ld32(lane_s + (uint32_t)(c * 32), r);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                unsigned long long z = 0ull;
                if (a.thresh) z = hash4(seed, a.salt, grp_row + (unsigned)(c * 8 + (j >> 2)));
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float e = ex2(fmaf(__uint_as_float(r[j + u]), scale2, kb[c * 32 + j + u]) - m);
                    sum += e;
                    const bool keep = a.thresh ? keep_of(z, u, a.thresh) : true;
                    r[j + u] = keep ? __float_as_uint(e) : 0u;
                }
            }
            tmem_st32(lane_s + (uint32_t)(c * 32), r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        fence_before();
        mbar_arrive(p_ready);
        red[half * kTile + rloc] = sum;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        sum = red[rloc] + red[kTile + rloc];
        if (half == 0 && blockIdx.z == 0 && grow < a.Tq && a.stats) {
            float2 *st = reinterpret_cast<float2 *>(a.stats) + ((size_t)bh * a.Tq + grow);
            *st = make_float2(m * kLn2, sum);                            // natural-log units: max of scale * s + bias
        }
        const float inv = a.inv_keep / sum;
        mbar_wait(o_full, 0);
        fence_after();
        // staging for coalesced stores: 32 rows x 36 floats per warp, aliasing the head of the pipeline ring (every MMA -
        // hence every shared-memory read - has retired when o_full fires and the producer has nothing left to load)
        float *stage_out = reinterpret_cast<float *>(smem) + (warp - 2) * (32 * 36);
        const int sub = lane & 7, rgrp = lane >> 3;
        const uint32_t lane_o = tmem_o + ((uint32_t)(q4 * 32) << 16);
        float *obase = a.out + (size_t)b * a.o_bs + col_base + dv0;
        const int noc = a.dv_tile / 32;
        const int o_lo = half ? (noc + 1) / 2 : 0, o_hi = half ? noc : (noc + 1) / 2;
#pragma unroll 1
        for (int c = o_lo; c < o_hi; ++c) {
            tmem_ld32(lane_o + (uint32_t)(c * 32), r);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4 *>(stage_out + lane * 36 + j) =
                    make_float4(__uint_as_float(r[j]) * inv, __uint_as_float(r[j + 1]) * inv,
                                __uint_as_float(r[j + 2]) * inv, __uint_as_float(r[j + 3]) * inv);
            __syncwarp();
            // row rr of this warp's 32 was scaled by ITS thread's `inv` before staging, so the transposed read is final
#pragma unroll
            for (int it8 = 0; it8 < 8; ++it8) {
                const int rr = it8 * 4 + rgrp;
                const float4 v = *reinterpret_cast<const float4 *>(stage_out + rr * 36 + sub * 4);
                const int orow = qt * kTile + q4 * 32 + rr;
                if (orow < a.Tq)
                    *reinterpret_cast<float4 *>(obase + (size_t)orow * a.o_ld + c * 32 + sub * 4) = v;
            }
            __syncwarp();
        }
        fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_s), "n"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
struct DsArgs {
    int B, H, Tq, Nk, D, Nk_pad, stages;
    float scale, inv_keep;
    unsigned thresh, salt;
    const float *key_bias, *out, *dout, *stats;
    long long o_ld, o_bs;
    float *ws_ds, *ws_p;
    const long long *seed_used;
};

__global__ void __launch_bounds__(kThreadsAttn, 1)
attn_bwd_ds_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_v, const DsArgs a)
{
    constexpr int kStage = 4 * kTileBytes;                               // Q | K_j | dO | V_j chunks, 16 KB each
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + a.stages * kStage);
    uint64_t *empty_bar = full_bar + a.stages;
    uint64_t *acc_full = empty_bar + a.stages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_full + 1);
    float *kb = reinterpret_cast<float *>(tmem_slot + 4);                 // [128]: log2e * key bias of this key chunk
    float *dl = kb + kTile;                                               // [128]: delta of this query tile's rows

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, bh = blockIdx.y, kc = blockIdx.z;
    const int b = bh / a.H, h = bh - b * a.H;
    const int col_base = h * a.D;
    const int nd = a.D / kChunk;
    const int key0 = kc * kTile;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tm_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tm_k) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tm_do) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tm_v) : "memory");
        for (int s = 0; s < a.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= 2 && warp < 6) {
        const int c = threadIdx.x - 64, key = key0 + c;
        kb[c] = key < a.Nk ? (a.key_bias ? kLog2e * __ldg(a.key_bias + (size_t)b * a.Nk + key) : 0.f) : -INFINITY;
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_s = *tmem_slot, tmem_dp = tmem_s + 128u;

    if (warp == 0) {
        if (lane == 0) {
            for (int dc = 0; dc < nd; ++dc) {
                const int s = dc % a.stages;
                mbar_wait(&empty_bar[s], ((dc / a.stages) & 1) ^ 1);
                mbar_expect_tx(&full_bar[s], kStage);
                uint8_t *st = smem + s * kStage;
                const int c0 = col_base + dc * kChunk;
                tma_load_3d(st, &tm_q, &full_bar[s], c0, qt * kTile, b);
                tma_load_3d(st + kTileBytes, &tm_k, &full_bar[s], c0, key0, b);
                tma_load_3d(st + 2 * kTileBytes, &tm_do, &full_bar[s], c0, qt * kTile, b);
                tma_load_3d(st + 3 * kTileBytes, &tm_v, &full_bar[s], c0, key0, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, false, false);
            for (int dc = 0; dc < nd; ++dc) {
                const int s = dc % a.stages;
                mbar_wait(&full_bar[s], (dc / a.stages) & 1);
                fence_after();
                const uint32_t sq = smem_u32(smem + s * kStage);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    umma_ss(tmem_s, desc_k_sw128(sq + k * 32), desc_k_sw128(sq + kTileBytes + k * 32), idesc,
                            (dc | k) != 0 ? 1u : 0u);
                    umma_ss(tmem_dp, desc_k_sw128(sq + 2 * kTileBytes + k * 32), desc_k_sw128(sq + 3 * kTileBytes + k * 32),
                            idesc, (dc | k) != 0 ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(acc_full);
        }
    } else {
        const int q4 = warp & 3;
        const int half = (warp - 2) >> 2;                                // key columns [64 * half, 64 * half + 64) of the chunk
        const int row0 = qt * kTile + q4 * 32;
        const int grow = row0 + lane;
        // delta[row] = <dO[row], O[row]> over the head's D columns (= sum_j P~ dP~).  The two warps of a lane quadrant take 16
        // of its 32 rows each, four rows per round with all of their (coalesced) loads in flight before the reductions; the
        // results meet in shared memory.
        for (int r0 = half * 16; r0 < half * 16 + 16; r0 += 4) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int row = row0 + r0 + u;
                const float *po = a.out + (size_t)b * a.o_bs + (size_t)row * a.o_ld + col_base;
                const float *pg = a.dout + (size_t)b * a.o_bs + (size_t)row * a.o_ld + col_base;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int d = lane + 32 * i;
                    if (d < a.D && row < a.Tq) acc[u] = fmaf(__ldg(po + d), __ldg(pg + d), acc[u]);
                }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
#pragma unroll
                for (int u = 0; u < 4; ++u) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
            }
            if (lane < 4) dl[q4 * 32 + r0 + lane] = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : (lane == 2 ? acc[2] : acc[3]));
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float delta = dl[q4 * 32 + lane];
        float m2 = 0.f, inv_l = 1.f;
        if (grow < a.Tq) {
            const float2 st = __ldg(reinterpret_cast<const float2 *>(a.stats) + ((size_t)bh * a.Tq + grow));
            m2 = st.x * kLog2e;
            inv_l = 1.f / st.y;
        }
        const unsigned long long seed = (a.thresh && a.seed_used) ? (unsigned long long)*a.seed_used : 0ull;
        const unsigned grp_row = (unsigned)((size_t)bh * a.Tq + grow) * (unsigned)(a.Nk_pad / 4);
        const float scale2 = a.scale * kLog2e;
        mbar_wait(acc_full, 0);
        fence_after();
        float *stage_ds = reinterpret_cast<float *>(smem) + (warp - 2) * (2 * 32 * 36);      // aliases the idle ring
        float *stage_p = stage_ds + 32 * 36;
        const int sub = lane & 7, rgrp = lane >> 3;
        const uint32_t lane_s = tmem_s + ((uint32_t)(q4 * 32) << 16), lane_dp = tmem_dp + ((uint32_t)(q4 * 32) << 16);
        uint32_t rs[32], rd[32];
#pragma unroll 1
        for (int c = 2 * half; c < 2 * half + 2; ++c) {
            tmem_ld32(lane_s + (uint32_t)(c * 32), rs);
            tmem_ld32(lane_dp + (uint32_t)(c * 32), rd);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float ds4[4], p4[4];
                unsigned long long z = 0ull;
                if (a.thresh) z = hash4(seed, a.salt, grp_row + (unsigned)((key0 + c * 32 + j) >> 2));
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float p = ex2(fmaf(__uint_as_float(rs[j + u]), scale2, kb[c * 32 + j + u]) - m2) * inv_l;
                    const bool keep = a.thresh ? keep_of(z, u, a.thresh) : true;
                    const float dp = keep ? __uint_as_float(rd[j + u]) * a.inv_keep : 0.f;
                    ds4[u] = a.scale * p * (dp - delta);
                    p4[u] = keep ? p * a.inv_keep : 0.f;
                }
                *reinterpret_cast<float4 *>(stage_ds + lane * 36 + j) = make_float4(ds4[0], ds4[1], ds4[2], ds4[3]);
                *reinterpret_cast<float4 *>(stage_p + lane * 36 + j) = make_float4(p4[0], p4[1], p4[2], p4[3]);
            }
            __syncwarp();
            const int col = key0 + c * 32 + sub * 4;
#pragma unroll
            for (int it8 = 0; it8 < 8; ++it8) {
                const int rr = it8 * 4 + rgrp;
                const int orow = row0 + rr;
                if (orow < a.Tq && col < a.Nk_pad) {
                    const size_t off = ((size_t)bh * a.Tq + orow) * (size_t)a.Nk_pad + col;
                    *reinterpret_cast<float4 *>(a.ws_ds + off) = *reinterpret_cast<const float4 *>(stage_ds + rr * 36 + sub * 4);
                    *reinterpret_cast<float4 *>(a.ws_p + off) = *reinterpret_cast<const float4 *>(stage_p + rr * 36 + sub * 4);
                }
            }
            __syncwarp();
        }
        fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_s), "n"(256) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The three gradient contractions of one attention call in ONE launch, operands read as stored (3-D TMA boxes):
//   problem 0   dQ[b, :, h] (+)= dS[bh] . K[b, :, h]       A = dS  K-major (contraction = keys),       B = K  MN-major
//   problem 1   dK[b, :, h] (+)= dS[bh]^T . Q[b, :, h]     A = dS  MN-major (contraction = query rows), B = Q  MN-major
//   problem 2   dV[b, :, h] (+)= P~[bh]^T . dO[b, :, h]    A = P~  MN-major,                            B = dO MN-major
// grid = (D / BLOCK_N, max row tiles, 3 * B * H); the A operands are the per-(batch, head) workspaces [B*H, Tq, Nk_pad],
// the B operands and the results head-sliced [B, T, H*D] tensors.
struct Bgemm3Args {
    int H, N, Z;
    int M[3], num_kb[3];
    float *C[3];
    long long c_ld[3], c_bs[3];
    int accumulate;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kThreads, 2)
attn_bgemm3_kernel(const __grid_constant__ CUtensorMap tm_a0, const __grid_constant__ CUtensorMap tm_b0,
                   const __grid_constant__ CUtensorMap tm_a1, const __grid_constant__ CUtensorMap tm_b1,
                   const __grid_constant__ CUtensorMap tm_a2, const __grid_constant__ CUtensorMap tm_b2, const Bgemm3Args g)
{
    constexpr int STAGES = 3;
    constexpr int kBBytes = BLOCK_N * kChunk * 4;
    constexpr int kStage = kTileBytes + kBBytes;
    const int prob = blockIdx.z / g.Z, z = blockIdx.z - prob * g.Z;
    const int n_blk = blockIdx.x, m_blk = blockIdx.y;
    // (selected with constant indices: a dynamically indexed kernel-parameter array would be copied to local memory)
    const int M = prob == 0 ? g.M[0] : (prob == 1 ? g.M[1] : g.M[2]);
    const int num_kb = prob == 0 ? g.num_kb[0] : (prob == 1 ? g.num_kb[1] : g.num_kb[2]);
    if (m_blk * kTile >= M) return;                                      // uniform per CTA, before any barrier / allocation
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + STAGES * kStage);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *acc_full = empty_bar + STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hz = z % g.H, zb = z / g.H;
    const int b_off = hz * g.N;
    const bool a_mn = prob != 0;
    const CUtensorMap *ta = prob == 0 ? &tm_a0 : (prob == 1 ? &tm_a1 : &tm_a2);
    const CUtensorMap *tb = prob == 0 ? &tm_b0 : (prob == 1 ? &tm_b1 : &tm_b2);

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(ta) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(tb) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(BLOCK_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&empty_bar[s], ((kb / STAGES) & 1) ^ 1);
                mbar_expect_tx(&full_bar[s], kStage);
                uint8_t *sa = smem + s * kStage;
                if (a_mn) {
#pragma unroll
                    for (int j = 0; j < kTile / 32; ++j)
                        tma_load_3d(sa + j * 4096, ta, &full_bar[s], m_blk * kTile + j * 32, kb * kChunk, z);
                } else {
                    tma_load_3d(sa, ta, &full_bar[s], kb * kChunk, m_blk * kTile, z);
                }
#pragma unroll
                for (int j = 0; j < BLOCK_N / 32; ++j)
                    tma_load_3d(sa + kTileBytes + j * 4096, tb, &full_bar[s], b_off + n_blk * BLOCK_N + j * 32, kb * kChunk, zb);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(BLOCK_N, a_mn, true);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&full_bar[s], (kb / STAGES) & 1);
                fence_after();
                const uint32_t sa = smem_u32(smem + s * kStage);
                const uint32_t sb = sa + kTileBytes;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t da = a_mn ? desc_mn_sw128(sa + k * 1024) : desc_k_sw128(sa + k * 32);
                    umma_ss(tmem_base, da, desc_mn_sw128(sb + k * 1024), idesc, (kb | k) != 0 ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(acc_full);
        }
    } else {
        const int q4 = warp & 3;
        mbar_wait(acc_full, 0);
        fence_after();
        float *stage_out = reinterpret_cast<float *>(smem) + (warp - 2) * (32 * 36);
        const int sub = lane & 7, rgrp = lane >> 3;
        float *cptr = prob == 0 ? g.C[0] : (prob == 1 ? g.C[1] : g.C[2]);
        const long long c_bs = prob == 0 ? g.c_bs[0] : (prob == 1 ? g.c_bs[1] : g.c_bs[2]);
        const long long c_ld = prob == 0 ? g.c_ld[0] : (prob == 1 ? g.c_ld[1] : g.c_ld[2]);
        float *cbase = cptr + (size_t)zb * c_bs + hz * g.N;
        uint32_t r[32];
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
            tmem_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(c * 32), r);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4 *>(stage_out + lane * 36 + j) =
                    make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                __uint_as_float(r[j + 3]));
            __syncwarp();
            const int col = n_blk * BLOCK_N + c * 32 + sub * 4;
#pragma unroll
            for (int it8 = 0; it8 < 8; ++it8) {
                const int rr = it8 * 4 + rgrp;
                const int grow = m_blk * kTile + q4 * 32 + rr;
                if (grow < M && col < g.N) {
                    const float4 v = *reinterpret_cast<const float4 *>(stage_out + rr * 36 + sub * 4);
                    float *dst = cbase + (size_t)grow * c_ld + col;
                    if (g.accumulate) red_add_v4(dst, v);
                    else *reinterpret_cast<float4 *>(dst) = v;
                }
            }
            __syncwarp();
        }
        fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(BLOCK_N) : "memory");
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// fp32 tensor [batches, rows, cols] (row stride ld, batch stride bs, in floats); box = [1, box_rows, 32 columns];
// zero fill outside [rows, cols] of each batch.  mn: the MN-major swizzle atom (32-byte granules), else the K-major 128B one
int make_map3(CUtensorMap *map, const float *ptr, uint64_t cols, uint64_t rows, uint64_t batches, uint64_t ld, uint64_t bs,
              uint32_t box_rows, bool mn) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return RLIPV2_ATTN_EDRIVER;
    if (batches <= 1 || bs == 0) bs = rows * ld;
    cuuint64_t dims[3] = {cols, rows, batches ? batches : 1};
    cuuint64_t strides[2] = {ld * sizeof(float), bs * sizeof(float)};
    cuuint32_t box[3] = {(cuuint32_t)kChunk, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, mn ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : RLIPV2_ATTN_EDRIVER;
}

inline bool aligned16(const void *p) { return (((uintptr_t)p) & 15) == 0; }
inline bool mult4(long long v) { return (v & 3) == 0; }

inline int key_pitch(int Nk) { return (Nk + 31) & ~31; }

// 16-bit dropout threshold (an element is kept when its 16 hash bits are >= it); 0 = no dropout
inline unsigned thresh16(double p) {
    if (p <= 0.0) return 0u;
    unsigned t = (unsigned)(p * 65536.0 + 0.5);
    return t < 1u ? 1u : (t > 65535u ? 65535u : t);
}

bool shape_ok(int B, int H, int Tq, int Nk, int D) {
    if (B <= 0 || H <= 0 || Tq <= 0 || Nk <= 0) return false;
    if (!(D == 32 || D == 64 || D == 128 || D == 256)) return false;
    return key_pitch(Nk) + 32 <= 512;
}

template <typename K>
int set_smem(K kern, int bytes) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    return e == cudaSuccess ? 0 : (int)e;
}

template <int BLOCK_N>
int launch_bgemm3(const CUtensorMap *t, const Bgemm3Args &g, cudaStream_t stream) {
    constexpr int smem = 3 * (kTileBytes + BLOCK_N * kChunk * 4) + 7 * 8 + 16 + 1024;
    auto kern = attn_bgemm3_kernel<BLOCK_N>;
    static bool configured = false;
    if (!configured) {
        int rc = set_smem(kern, smem);
        if (rc) return rc;
        configured = true;
    }
    int mt = 0;
    for (int p = 0; p < 3; ++p) mt = (g.M[p] + kTile - 1) / kTile > mt ? (g.M[p] + kTile - 1) / kTile : mt;
    dim3 grid((g.N + BLOCK_N - 1) / BLOCK_N, mt, 3 * g.Z);
    kern<<<grid, kThreads, smem, stream>>>(t[0], t[1], t[2], t[3], t[4], t[5], g);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

}  // namespace

extern "C" {

int rlipv2_attn_supported(int B, int H, int Tq, int Nk, int D) { return shape_ok(B, H, Tq, Nk, D) ? 1 : 0; }

int rlipv2_attn_key_pitch(int Nk) { return key_pitch(Nk); }

int rlipv2_attn_forward_tf32(const float *q, long long q_ld, long long q_bs, const float *k, long long k_ld, long long k_bs,
                             const float *v, long long v_ld, long long v_bs, const float *key_bias, float *out,
                             long long o_ld, long long o_bs, float *stats, int B, int H, int Tq, int Nk, int D, float scale,
                             double dropout_p, const long long *seed, unsigned salt, long long *seed_used, void *stream)
{
    if (!q || !k || !v || !out || !stats || dropout_p < 0.0 || dropout_p >= 1.0) return RLIPV2_ATTN_EINVAL;
    if (!shape_ok(B, H, Tq, Nk, D)) return RLIPV2_ATTN_ESHAPE;
    if (!mult4(q_ld) || !mult4(q_bs) || !mult4(k_ld) || !mult4(k_bs) || !mult4(v_ld) || !mult4(v_bs) || !mult4(o_ld) ||
        !mult4(o_bs))
        return RLIPV2_ATTN_ESHAPE;
    if (!aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(out) || (((uintptr_t)stats) & 7)) return RLIPV2_ATTN_EALIGN;
    FwdArgs a;
    a.B = B; a.H = H; a.Tq = Tq; a.Nk = Nk; a.D = D;
    a.Nk_pad = key_pitch(Nk);
    a.dv_tile = D;
    while (a.Nk_pad + a.dv_tile > 512) a.dv_tile >>= 1;                  // >= 32 by shape_ok
    a.nbox = (a.Nk_pad + 255) / 256;
    a.box_rows = a.Nk_pad / a.nbox;                                      // Nk_pad % 32 == 0, nbox <= 2: a multiple of 16
    const int st_a = kTileBytes + a.Nk_pad * 128, st_b = a.dv_tile * 128;
    a.stage_bytes = st_a > st_b ? st_a : st_b;
    const int total = D / kChunk + a.Nk_pad / kChunk;
    a.stages = (200 * 1024) / a.stage_bytes;
    if (a.stages > 4) a.stages = 4;
    if (a.stages > total) a.stages = total;
    if (a.stages < 2) return RLIPV2_ATTN_ESHAPE;
    a.scale = scale;
    a.thresh = thresh16(dropout_p);
    a.inv_keep = dropout_p > 0.0 ? (float)(1.0 / (1.0 - dropout_p)) : 1.f;
    a.salt = salt;
    a.key_bias = key_bias;
    a.out = out; a.o_ld = o_ld; a.o_bs = o_bs;
    a.stats = stats;
    a.seed = seed; a.seed_used = seed_used;
    CUtensorMap tq, tk, tv;
    int rc = make_map3(&tq, q, (uint64_t)H * D, (uint64_t)Tq, (uint64_t)B, (uint64_t)q_ld, (uint64_t)q_bs, kTile, false);
    if (rc) return rc;
    rc = make_map3(&tk, k, (uint64_t)H * D, (uint64_t)Nk, (uint64_t)B, (uint64_t)k_ld, (uint64_t)k_bs, (uint32_t)a.box_rows, false);
    if (rc) return rc;
    rc = make_map3(&tv, v, (uint64_t)H * D, (uint64_t)Nk, (uint64_t)B, (uint64_t)v_ld, (uint64_t)v_bs, 32, true);
    if (rc) return rc;
    const int smem = a.stages * a.stage_bytes + (2 * a.stages + 3) * 8 + 16 + (2 * kTile + a.Nk_pad) * 4 + 1024;
    static int configured = 0;
    if (smem > configured) {
        rc = set_smem(attn_fwd_kernel, smem);
        if (rc) return rc;
        configured = smem;
    }
    dim3 grid((Tq + kTile - 1) / kTile, B * H, D / a.dv_tile);
    attn_fwd_kernel<<<grid, kThreadsAttn, smem, (cudaStream_t)stream>>>(tq, tk, tv, a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

int rlipv2_attn_backward_tf32(const float *q, long long q_ld, long long q_bs, const float *k, long long k_ld, long long k_bs,
                              const float *v, long long v_ld, long long v_bs, const float *key_bias, const float *out,
                              const float *dout, long long o_ld, long long o_bs, const float *stats, float *dq,
                              long long dq_ld, long long dq_bs, float *dk, long long dk_ld, long long dk_bs, float *dv,
                              long long dv_ld, long long dv_bs, float *ws_ds, float *ws_p, int B, int H, int Tq, int Nk,
                              int D, float scale, double dropout_p, const long long *seed_used, unsigned salt,
                              int accumulate, void *stream)
{
    if (!q || !k || !v || !out || !dout || !stats || !dq || !dk || !dv || !ws_ds || !ws_p || dropout_p < 0.0 ||
        dropout_p >= 1.0)
        return RLIPV2_ATTN_EINVAL;
    if (!shape_ok(B, H, Tq, Nk, D)) return RLIPV2_ATTN_ESHAPE;
    const long long strides[] = {q_ld, q_bs, k_ld, k_bs, v_ld, v_bs, o_ld, o_bs, dq_ld, dq_bs, dk_ld, dk_bs, dv_ld, dv_bs};
    for (long long s : strides)
        if (!mult4(s)) return RLIPV2_ATTN_ESHAPE;
    if (!aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(out) || !aligned16(dout) || !aligned16(dq) ||
        !aligned16(dk) || !aligned16(dv) || !aligned16(ws_ds) || !aligned16(ws_p) || (((uintptr_t)stats) & 7))
        return RLIPV2_ATTN_EALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    const int Nk_pad = key_pitch(Nk), Z = B * H;
    const uint64_t HD = (uint64_t)H * D;
    // 1. dS and the dropped probabilities, one CTA per (query tile, batch x head, 128-key chunk)
    {
        DsArgs a;
        a.B = B; a.H = H; a.Tq = Tq; a.Nk = Nk; a.D = D; a.Nk_pad = Nk_pad;
        a.stages = D / kChunk < 3 ? (D / kChunk < 2 ? 2 : D / kChunk) : 3;     // >= 2: the epilogue stages through the ring
        a.scale = scale;
        a.thresh = thresh16(dropout_p);
        a.inv_keep = dropout_p > 0.0 ? (float)(1.0 / (1.0 - dropout_p)) : 1.f;
        a.salt = salt;
        a.key_bias = key_bias; a.out = out; a.dout = dout; a.stats = stats;
        a.o_ld = o_ld; a.o_bs = o_bs;
        a.ws_ds = ws_ds; a.ws_p = ws_p;
        a.seed_used = seed_used;
        CUtensorMap tq, tk, tg, tv;
        int rc = make_map3(&tq, q, HD, (uint64_t)Tq, (uint64_t)B, (uint64_t)q_ld, (uint64_t)q_bs, kTile, false);
        if (rc) return rc;
        rc = make_map3(&tk, k, HD, (uint64_t)Nk, (uint64_t)B, (uint64_t)k_ld, (uint64_t)k_bs, kTile, false);
        if (rc) return rc;
        rc = make_map3(&tg, dout, HD, (uint64_t)Tq, (uint64_t)B, (uint64_t)o_ld, (uint64_t)o_bs, kTile, false);
        if (rc) return rc;
        rc = make_map3(&tv, v, HD, (uint64_t)Nk, (uint64_t)B, (uint64_t)v_ld, (uint64_t)v_bs, kTile, false);
        if (rc) return rc;
        const int smem = a.stages * 4 * kTileBytes + (2 * a.stages + 1) * 8 + 16 + 2 * 128 * 4 + 1024;
        static int configured = 0;
        if (smem > configured) {
            rc = set_smem(attn_bwd_ds_kernel, smem);
            if (rc) return rc;
            configured = smem;
        }
        dim3 grid((Tq + kTile - 1) / kTile, Z, (Nk + kTile - 1) / kTile);
        attn_bwd_ds_kernel<<<grid, kThreadsAttn, smem, s>>>(tq, tk, tg, tv, a);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    const int block_n = D >= 128 ? 128 : D;
    CUtensorMap t[6];
    // problem 0: dS K-major, K as stored;  1: dS MN-major, Q as stored;  2: P~ MN-major, dO as stored
    int rc = make_map3(&t[0], ws_ds, (uint64_t)Nk_pad, (uint64_t)Tq, (uint64_t)Z, (uint64_t)Nk_pad, (uint64_t)Tq * Nk_pad, kTile, false);
    if (rc) return rc;
    rc = make_map3(&t[1], k, HD, (uint64_t)Nk, (uint64_t)B, (uint64_t)k_ld, (uint64_t)k_bs, 32, true);
    if (rc) return rc;
    rc = make_map3(&t[2], ws_ds, (uint64_t)Nk_pad, (uint64_t)Tq, (uint64_t)Z, (uint64_t)Nk_pad, (uint64_t)Tq * Nk_pad, 32, true);
    if (rc) return rc;
    rc = make_map3(&t[3], q, HD, (uint64_t)Tq, (uint64_t)B, (uint64_t)q_ld, (uint64_t)q_bs, 32, true);
    if (rc) return rc;
    rc = make_map3(&t[4], ws_p, (uint64_t)Nk_pad, (uint64_t)Tq, (uint64_t)Z, (uint64_t)Nk_pad, (uint64_t)Tq * Nk_pad, 32, true);
    if (rc) return rc;
    rc = make_map3(&t[5], dout, HD, (uint64_t)Tq, (uint64_t)B, (uint64_t)o_ld, (uint64_t)o_bs, 32, true);
    if (rc) return rc;
    Bgemm3Args g;
    g.H = H; g.N = D; g.Z = Z; g.accumulate = accumulate;
    g.M[0] = Tq; g.num_kb[0] = Nk_pad / kChunk; g.C[0] = dq; g.c_ld[0] = dq_ld; g.c_bs[0] = dq_bs;
    g.M[1] = Nk; g.num_kb[1] = (Tq + kChunk - 1) / kChunk; g.C[1] = dk; g.c_ld[1] = dk_ld; g.c_bs[1] = dk_bs;
    g.M[2] = Nk; g.num_kb[2] = (Tq + kChunk - 1) / kChunk; g.C[2] = dv; g.c_ld[2] = dv_ld; g.c_bs[2] = dv_bs;
    switch (block_n) {
        case 32: return launch_bgemm3<32>(t, g, s);
        case 64: return launch_bgemm3<64>(t, g, s);
        default: return launch_bgemm3<128>(t, g, s);
    }
}

const char *rlipv2_attn_error_string(int code)
{
    switch (code) {
        case 0: return "success";
        case RLIPV2_ATTN_EINVAL: return "rlipv2_attn: invalid argument";
        case RLIPV2_ATTN_ESHAPE: return "rlipv2_attn: shape not supported (D in {32,64,128,256}, Nk <= 480, strides multiples of 4 floats)";
        case RLIPV2_ATTN_EALIGN: return "rlipv2_attn: pointers must be 16-byte aligned";
        case RLIPV2_ATTN_EDRIVER: return "rlipv2_attn: cuTensorMapEncodeTiled unavailable or failed";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "rlipv2_attn: unknown error";
    }
}

unsigned long long rlipv2_attn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
