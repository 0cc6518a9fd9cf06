// dense_tf32.cu - y = act(x . W^T + bias) on the 5th-generation tensor cores (tcgen05), sm_100a.
//
// The dense contractions of the ParSeDA hot path are `nn.Linear`s on fp32 activations:
//   ALIF projections   /root/reference/models/fuse_helper.py:370-373, 463-464  (256|768 <-> 2048)
//   RobertaLayer       models/modeling_roberta.py:159-235, 252-256, 318-336    (768 <-> 768|3072)
//   deformable FFN     models/dab_deformable/deformable_transformer.py:1283-1287 (256 <-> 2048)
//   MSDeformAttn proj  models/ops/modules/ms_deform_attn.py:98-118             (256 -> 256|128)
// x is [M, K] row-major, W is nn.Linear's [N, K] row-major: both operands are K-major, the native
// layout of tcgen05.mma.  Products are TF32 (10-bit mantissa, what the reference's pinned torch 1.10
// computes by default on tensor-core GPUs), accumulation is fp32 in tensor memory.
//
// Kernel structure (one 128 x BLOCK_N output tile per CTA, 6 warps):
//   warp 0 / lane 0   TMA producer: cp.async.bulk.tensor 2-D loads of a 128x32 (A) and BLOCK_Nx32 (B)
//                     fp32 box per stage into 128B-swizzled shared memory, mbarrier expect_tx
//   warp 1            allocates BLOCK_N TMEM columns; lane 0 issues 4 x tcgen05.mma.kind::tf32
//                     (M=128, N=BLOCK_N, K=8) per stage, tcgen05.commit frees the stage / signals
//                     the epilogue
//   warps 2-5         epilogue: tcgen05.ld 32 lanes x 32 columns at a time (warp w owns TMEM lane
//                     quadrant w % 4), + bias, ReLU / exact GELU, transposed through shared memory
//                     so that every global store instruction writes whole 128-byte row segments
// STAGES-deep full/empty mbarrier ring between producer and MMA issuer; 3 stages (96 KB) so that two
// CTAs share an SM and one CTA's epilogue overlaps the other's main loop.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "rlipv2_dense.h"

namespace {

std::atomic<unsigned long long> g_launches{0};
// small-grid tuning of the forward linear (rlipv2_dense_set_small_mode): 0 = always 128x128 tiles, 3 stages, 2 CTAs/SM;
// 1 = grids of at most one CTA per SM run a 6-stage ring (1 CTA/SM); 2 = additionally 128x64 tiles with an 8-stage ring
// while that keeps the grid within one CTA per SM
std::atomic<int> g_small_mode{2};
// grids of more than this many 128 x 128 tiles run the persistent kernel (0 = never); rlipv2_dense_set_persistent_min_tiles
std::atomic<int> g_persistent_min_tiles{0};
// the gated input gradient (encoder FFN backward) on its persistent kernel too (rlipv2_dense_set_persistent_dgrad)
std::atomic<int> g_persistent_dgrad{0};
constexpr int kNumSMs = 148;

constexpr int kBlockM = 128;
constexpr int kBlockK = 32;                 // 32 fp32 = 128 bytes = one swizzle-128B row
constexpr int kUmmaK = 8;                   // tf32: 32 bytes per MMA K-step
constexpr int kThreads = 192;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    // try_wait suspends for a hardware-defined time slice per attempt; a barrier that never flips (a byte-count
    // or descriptor bug) must become an error, not a hung GPU: trap after ~2^26 attempts (seconds)
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (spins > (1u << 26)) __trap();
    }
}

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :: "r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// K-major, 128-byte swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart
// (cute::UMMA::SmemDescriptor: start>>4 | LBO(1)<<16 | SBO(1024>>4)<<32 | version 1<<46 | SWIZZLE_128B(2)<<61)
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// MN-major operand tile (the operand is stored [K rows, MN cols], MN contiguous - the layout of both operands of
// a weight-gradient GEMM and of the weight in an input-gradient GEMM).  For 32-bit (TF32) MN-major operands the
// only shared-memory layout tcgen05 accepts is SWIZZLE_128B_BASE32B (cute: Layout_MN_SW128_32B_Atom, Swizzle<2,5,2>
// - 32-byte granules of a 128-byte row XOR-ed with the row index mod 4; TMA writes it with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).  Shared memory holds BLOCK/32 chunks of [32 K rows x 32 MN elements
// (128 B)], 4096 B apart; one MMA (K = 8) reads 8 consecutive 128-byte rows = two 4-row swizzle atoms of every
// chunk.  Canonical form (16-byte units): ((8,n),(4,k)) : ((1,LBO),(8,SBO)) with LBO = chunk stride = 4096 B,
// SBO = 4-row group stride = 512 B; layout type 1 in bits [61,64).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}

__device__ __forceinline__ void red_add_v4(float *addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

template <int BLOCK_N>
struct SmemLayout {
    static constexpr int kABytes = kBlockM * kBlockK * 4;        // 16 KB
    static constexpr int kBBytes = BLOCK_N * kBlockK * 4;
    static constexpr int kStageBytes = kABytes + kBBytes;
};

// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1)<<4, a/b format TF32 (2)<<7 / <<10,
// a/b K-major (0), n_dim = N>>3 at bit 17, m_dim = M>>4 at bit 24
template <int BLOCK_N>
__host__ __device__ constexpr uint32_t make_idesc() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
}

// same with the operand majors: bit 15 = A is MN-major, bit 16 = B is MN-major
template <int BLOCK_N, bool A_MN, bool B_MN>
__host__ __device__ constexpr uint32_t make_idesc_major() {
    return make_idesc<BLOCK_N>() | (A_MN ? (1u << 15) : 0u) | (B_MN ? (1u << 16) : 0u);
}

// ---------------------------------------------------------------------------------------------------------------
// General TF32 GEMM for the backward contractions:  C[M,N] (op)= A . B^T  with logical A [M,K], B [N,K] and
//   A_MN: A is stored [K, M] (M contiguous)      B_MN: B is stored [K, N] (N contiguous)
// EPI_STORE   C = acc
// EPI_ATOMIC  C += acc with red.global.add.v4.f32 (split-K over blockIdx.z; C zero-initialised by the caller)
// EPI_MASK    C = (mask > 0) ? acc : 0 and colsum[m_blk, n] = sum over the CTA's 128 rows of C (the ReLU backward +
//             bias gradient of the layer that produced `mask`, fused into the input-gradient GEMM; the caller
//             sums the gridDim.y row-tile partials)
// Same warp roles / pipeline as linear_tf32_kernel.  K need not be a multiple of 32 for MN-major operands (TMA
// zero-fills rows past the end); M and N tails are handled by TMA zero fill + guarded stores.
// ---------------------------------------------------------------------------------------------------------------
// EPI_ATOMIC_BIAS  EPI_ATOMIC for a split-K *forward* linear: the blockIdx.z == 0 slice also adds bias[n] (passed in the
//             `mask` argument) to its partial, so that zero-initialised C ends up as A . B^T + bias
enum { EPI_STORE = 0, EPI_ATOMIC = 1, EPI_MASK = 2, EPI_ATOMIC_BIAS = 3 };

template <int BLOCK_N, int STAGES, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(kThreads, 2)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 float *__restrict__ C, const float *__restrict__ mask, float *__restrict__ colsum,
                 int M, int N, int num_kb, int kb_per_split)
{
    using L = SmemLayout<BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + STAGES * L::kStageBytes);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tmem_full_bar = empty_bar + STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_blk = blockIdx.x, m_blk = blockIdx.y;
    const int kb0 = blockIdx.z * kb_per_split;
    const int kb1 = min(num_kb, kb0 + kb_per_split);
    const int nk = kb1 - kb0;                       // >= 1 by construction of the grid

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tm_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tm_b) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(BLOCK_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                                   // ---- TMA producer
            for (int i = 0; i < nk; ++i) {
                const int kb = kb0 + i;
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_expect_tx(&full_bar[s], L::kStageBytes);
                uint8_t *sa = smem + s * L::kStageBytes;
                if (A_MN) {                                // box = [32 k rows][32 m], one per 32-wide chunk of M
#pragma unroll
                    for (int j = 0; j < kBlockM / 32; ++j)
                        tma_load_2d(sa + j * 4096, &tm_a, &full_bar[s], m_blk * kBlockM + j * 32, kb * kBlockK);
                } else {
                    tma_load_2d(sa, &tm_a, &full_bar[s], kb * kBlockK, m_blk * kBlockM);
                }
                if (B_MN) {
#pragma unroll
                    for (int j = 0; j < BLOCK_N / 32; ++j)
                        tma_load_2d(sa + L::kABytes + j * 4096, &tm_b, &full_bar[s], n_blk * BLOCK_N + j * 32, kb * kBlockK);
                } else {
                    tma_load_2d(sa + L::kABytes, &tm_b, &full_bar[s], kb * kBlockK, n_blk * BLOCK_N);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                   // ---- MMA issuer
            constexpr uint32_t idesc = make_idesc_major<BLOCK_N, A_MN, B_MN>();
            for (int i = 0; i < nk; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(smem + s * L::kStageBytes);
                const uint32_t sb = sa + L::kABytes;
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                    // K-major: +32 bytes inside the swizzle atom; MN-major: +8 rows of 128 bytes
                    const uint64_t da = A_MN ? umma_desc_mn_sw128(sa + k * 1024) : umma_desc_k_sw128(sa + k * kUmmaK * 4);
                    const uint64_t db = B_MN ? umma_desc_mn_sw128(sb + k * 1024) : umma_desc_k_sw128(sb + k * kUmmaK * 4);
                    umma_tf32(tmem_base, da, db, idesc, (i | k) != 0 ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(tmem_full_bar);
        }
    } else {                                               // ---- epilogue warps 2..5
        const int q = warp & 3;
        float *stage_out = reinterpret_cast<float *>(smem) + (warp - 2) * (32 * 36);
        float *col_part = reinterpret_cast<float *>(smem) + 4 * (32 * 36);        // [4 warps][BLOCK_N], EPI_MASK only
        const int sub = lane & 7, rgrp = lane >> 3;
        const size_t col0 = (size_t)n_blk * BLOCK_N;
        // EPI_MASK: the 8 mask vectors this lane needs for a 32-column chunk are fetched one chunk ahead (the
        // first before the accumulator is even complete), so the epilogue pays one memory latency per chunk at
        // most instead of one per row group
        float4 mk[8];
        auto fetch_mask = [&](int c) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int grow = m_blk * kBlockM + q * 32 + it * 4 + rgrp;
                const size_t col = col0 + c * 32 + sub * 4;
                mk[it] = (grow < M && col < (size_t)N) ? __ldg(reinterpret_cast<const float4 *>(mask + (size_t)grow * N + col))
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        if (EPI == EPI_MASK) fetch_mask(0);
        mbar_wait(tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
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
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4 *>(stage_out + lane * 36 + j) =
                    make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                __uint_as_float(r[j + 3]));
            __syncwarp();
            const size_t col = col0 + c * 32 + sub * 4;
            float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int rr = it * 4 + rgrp;
                float4 v = *reinterpret_cast<const float4 *>(stage_out + rr * 36 + sub * 4);
                const int grow = m_blk * kBlockM + q * 32 + rr;
                if (grow < M && col < (size_t)N) {
                    float *dst = C + (size_t)grow * N + col;
                    if (EPI == EPI_ATOMIC) {
                        red_add_v4(dst, v);
                    } else if (EPI == EPI_ATOMIC_BIAS) {
                        if (blockIdx.z == 0 && mask != nullptr) {
                            const float4 bb = __ldg(reinterpret_cast<const float4 *>(mask + col));
                            v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
                        }
                        red_add_v4(dst, v);
                    } else {
                        if (EPI == EPI_MASK) {
                            v.x = mk[it].x > 0.f ? v.x : 0.f; v.y = mk[it].y > 0.f ? v.y : 0.f;
                            v.z = mk[it].z > 0.f ? v.z : 0.f; v.w = mk[it].w > 0.f ? v.w : 0.f;
                            csum.x += v.x; csum.y += v.y; csum.z += v.z; csum.w += v.w;
                        }
                        *reinterpret_cast<float4 *>(dst) = v;
                    }
                }
            }
            if (EPI == EPI_MASK) {
                // lanes with equal `sub` hold the same 4 columns: fold the 4 row groups of this warp's 32 rows
#pragma unroll
                for (int o = 8; o < 32; o <<= 1) {
                    csum.x += __shfl_xor_sync(0xffffffffu, csum.x, o); csum.y += __shfl_xor_sync(0xffffffffu, csum.y, o);
                    csum.z += __shfl_xor_sync(0xffffffffu, csum.z, o); csum.w += __shfl_xor_sync(0xffffffffu, csum.w, o);
                }
                if (rgrp == 0)
                    *reinterpret_cast<float4 *>(col_part + (warp - 2) * BLOCK_N + c * 32 + sub * 4) = csum;
                if (c + 1 < BLOCK_N / 32) fetch_mask(c + 1);
            }
            __syncwarp();
        }
        if (EPI == EPI_MASK) {
            // the CTA's 128 rows: sum the 4 warps' partials, one row of colsum[gridDim.y, N] per CTA row-tile
            // (the caller reduces the row-tiles; no atomics, deterministic)
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int t = threadIdx.x - 64;                    // 0..127 over the epilogue warps
            for (int cc = t; cc < BLOCK_N; cc += 128) {
                const float v = col_part[cc] + col_part[BLOCK_N + cc] + col_part[2 * BLOCK_N + cc] + col_part[3 * BLOCK_N + cc];
                if (col0 + cc < (size_t)N) colsum[(size_t)m_blk * N + col0 + cc] = v;
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(BLOCK_N) : "memory");
    }
}

template <int BLOCK_N, int STAGES, int ACT>
__global__ void __launch_bounds__(kThreads, 2)
linear_tf32_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                   const float *__restrict__ bias, const uint8_t *__restrict__ rowmask, float *__restrict__ C,
                   int M, int N, int K)
{
    using L = SmemLayout<BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    // 128B-swizzled TMA/UMMA tiles need a 1024-byte aligned base (aligned in the shared window)
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + STAGES * L::kStageBytes);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tmem_full_bar = empty_bar + STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_blk = blockIdx.x, m_blk = blockIdx.y;
    const int num_k = K / kBlockK;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tm_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tm_b) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {      // whole warp: allocate the accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(BLOCK_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                                   // ---- TMA producer
            for (int kb = 0; kb < num_k; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_expect_tx(&full_bar[s], L::kStageBytes);
                uint8_t *sa = smem + s * L::kStageBytes;
                tma_load_2d(sa, &tm_a, &full_bar[s], kb * kBlockK, m_blk * kBlockM);
                tma_load_2d(sa + L::kABytes, &tm_b, &full_bar[s], kb * kBlockK, n_blk * BLOCK_N);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                   // ---- MMA issuer
            constexpr uint32_t idesc = make_idesc<BLOCK_N>();
            for (int kb = 0; kb < num_k; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(smem + s * L::kStageBytes);
                const uint32_t sb = sa + L::kABytes;
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                    // advancing K inside the swizzle atom = advancing the start address by 32 bytes
                    umma_tf32(tmem_base, umma_desc_k_sw128(sa + k * kUmmaK * 4), umma_desc_k_sw128(sb + k * kUmmaK * 4),
                              idesc, (kb | k) != 0 ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);                // stage reusable once these MMAs retire
            }
            umma_commit(tmem_full_bar);                    // accumulator complete
        }
    } else {                                               // ---- epilogue warps 2..5
        mbar_wait(tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;                            // TMEM lane quadrant this warp may read
        // Staging area for coalesced stores: 32 rows x 36 floats per warp.  It aliases pipeline stage
        // 0, which is idle by now: tmem_full_bar fires only after every MMA (hence every smem read)
        // has retired and the producer has no load left to issue.
        float *stage_out = reinterpret_cast<float *>(smem) + (warp - 2) * (32 * 36);
        const int sub = lane & 7, rgrp = lane >> 3;
        const size_t col0 = (size_t)n_blk * BLOCK_N;
        const float *brow = bias ? bias + col0 : nullptr;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
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
            // thread `lane` holds row (q*32 + lane), columns c*32 .. c*32+31: bias + activation, then
            // transpose through shared memory so that 8 lanes write one 128-byte row segment
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 o = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                       __uint_as_float(r[j + 3]));
                if (brow) {
                    const float4 b = __ldg(reinterpret_cast<const float4 *>(brow + c * 32 + j));
                    o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                }
                if (ACT == RLIPV2_DENSE_ACT_RELU) {
                    o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                } else if (ACT == RLIPV2_DENSE_ACT_GELU) {
                    o.x = gelu_exact(o.x); o.y = gelu_exact(o.y); o.z = gelu_exact(o.z); o.w = gelu_exact(o.w);
                }
                *reinterpret_cast<float4 *>(stage_out + lane * 36 + j) = o;
            }
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int rr = it * 4 + rgrp;
                const float4 v = *reinterpret_cast<const float4 *>(stage_out + rr * 36 + sub * 4);
                const int grow = m_blk * kBlockM + q * 32 + rr;
                if (grow < M) {
                    // rowmask: `value.masked_fill(padding_mask[..., None], 0)` of ms_deform_attn.py:99-100 in place
                    const float4 o = (rowmask && rowmask[grow]) ? make_float4(0.f, 0.f, 0.f, 0.f) : v;
                    *reinterpret_cast<float4 *>(C + (size_t)grow * N + col0 + c * 32 + sub * 4) = o;
                }
            }
            __syncwarp();
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(BLOCK_N) : "memory");
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Persistent variant of linear_tf32_kernel for large grids (the encoder's 44k-row linears: thousands of 128 x 128
// tiles with only 8 k-blocks each at K = 256).  One CTA per SM walks tiles b, b + G, b + 2G, ... (n fastest, so that
// CTAs running at the same time share the rows of x in L2); the TMA ring runs on across tile boundaries and the
// accumulator is double-buffered in tensor memory (2 x 128 columns): while the epilogue warps drain tile i, the
// producer / MMA warps are already on tile i + 1.  The one-tile-per-CTA kernel pays barrier init, TMEM allocation,
// pipeline fill and an un-overlapped epilogue per tile - more than the 8 k-blocks of math themselves.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive_plain(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}

template <int BLOCK_N, int STAGES, int ACT>
__global__ void __launch_bounds__(kThreads, 1)
linear_tf32_persistent_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                              const float *__restrict__ bias, const uint8_t *__restrict__ rowmask, float *__restrict__ C,
                              int M, int N, int K)
{
    // BLOCK_N = 256 (N % 256 == 0): a 128 x 256 tile fetches 25 % fewer operand bytes per output than two 128 x 128 tiles -
    // with every SM pulling operands at once the kernel is bound by the L2 -> SM bandwidth (~81 GB/s per SM: measured
    // 132.8 us for the encoder FFN-up against 119 us of operand traffic at that rate); 2 x 256 accumulator columns = all of TMEM
    using L = SmemLayout<BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float *stage_all = reinterpret_cast<float *>(smem + STAGES * L::kStageBytes);          // 4 warps x 32 x 36 floats
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(stage_all + 4 * 32 * 36);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tmem_full = empty_bar + STAGES;          // [2]
    uint64_t *tmem_empty = tmem_full + 2;              // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_k = K / kBlockK;
    const int n_tiles = N / BLOCK_N;
    const int num_tiles = n_tiles * ((M + kBlockM - 1) / kBlockM);

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tm_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tm_b) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(2 * BLOCK_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                                   // ---- TMA producer: the ring does not stop at tile boundaries
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
                for (int kb = 0; kb < num_k; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                    mbar_expect_tx(&full_bar[s], L::kStageBytes);
                    uint8_t *sa = smem + s * L::kStageBytes;
                    tma_load_2d(sa, &tm_a, &full_bar[s], kb * kBlockK, m_blk * kBlockM);
                    tma_load_2d(sa + L::kABytes, &tm_b, &full_bar[s], kb * kBlockK, n_blk * BLOCK_N);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                   // ---- MMA issuer
            constexpr uint32_t idesc = make_idesc<BLOCK_N>();
            uint32_t it = 0, local = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
                const uint32_t acc = local & 1;
                mbar_wait(&tmem_empty[acc], ((local >> 1) & 1) ^ 1);      // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tacc = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < num_k; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&full_bar[s], (it / STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = smem_u32(smem + s * L::kStageBytes);
                    const uint32_t sb = sa + L::kABytes;
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k)
                        umma_tf32(tacc, umma_desc_k_sw128(sa + k * kUmmaK * 4), umma_desc_k_sw128(sb + k * kUmmaK * 4), idesc,
                                  (kb | k) != 0 ? 1u : 0u);
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&tmem_full[acc]);
            }
        }
    } else {                                               // ---- epilogue warps 2..5
        const int q = warp & 3;
        float *stage_out = stage_all + (warp - 2) * (32 * 36);
        const int sub = lane & 7, rgrp = lane >> 3;
        uint32_t local = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
            const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
            const uint32_t acc = local & 1;
            mbar_wait(&tmem_full[acc], (local >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const size_t col0 = (size_t)n_blk * BLOCK_N;
            const float *brow = bias ? bias + col0 : nullptr;
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + acc * BLOCK_N + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
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
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 o = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                           __uint_as_float(r[j + 3]));
                    if (brow) {
                        const float4 b = __ldg(reinterpret_cast<const float4 *>(brow + c * 32 + j));
                        o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                    }
                    if (ACT == RLIPV2_DENSE_ACT_RELU) {
                        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                    } else if (ACT == RLIPV2_DENSE_ACT_GELU) {
                        o.x = gelu_exact(o.x); o.y = gelu_exact(o.y); o.z = gelu_exact(o.z); o.w = gelu_exact(o.w);
                    }
                    *reinterpret_cast<float4 *>(stage_out + lane * 36 + j) = o;
                }
                __syncwarp();
#pragma unroll
                for (int it8 = 0; it8 < 8; ++it8) {
                    const int rr = it8 * 4 + rgrp;
                    const float4 v = *reinterpret_cast<const float4 *>(stage_out + rr * 36 + sub * 4);
                    const int grow = m_blk * kBlockM + q * 32 + rr;
                    if (grow < M) {
                        const float4 o = (rowmask && rowmask[grow]) ? make_float4(0.f, 0.f, 0.f, 0.f) : v;
                        *reinterpret_cast<float4 *>(C + (size_t)grow * N + col0 + c * 32 + sub * 4) = o;
                    }
                }
                __syncwarp();
            }
            // this thread's TMEM reads of the accumulator are complete (wait::ld above): hand it back to the MMA warp
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive_plain(&tmem_empty[acc]);
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(2 * BLOCK_N) : "memory");
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Persistent input-gradient GEMM with the ReLU gate:  dx[T, N] = (mask > 0) ? g[T, Kc] . w[Kc, N] : 0  and per-row-tile column
// sums of the gated dx (EPI_MASK of gemm_tf32_kernel) - the encoder FFN's `gh = (g W2) * (h > 0)`, `db1 = colsum(gh)`:
// T = 44 446, N = 2048, Kc = 256, 364 MB of gate in and 364 MB of gradient out per call.  Same schedule as
// linear_tf32_persistent_kernel (one CTA per SM, 128 x 256 tiles, TMA ring across tiles, double-buffered TMEM accumulator);
// A = g K-major, B = w as stored (MN-major 32 x 32 boxes).
// ---------------------------------------------------------------------------------------------------------------
template <int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
dgrad_mask_persistent_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                             float *__restrict__ C, const float *__restrict__ mask, float *__restrict__ colsum,
                             int M, int N, int num_kb)
{
    constexpr int BLOCK_N = 256;
    using L = SmemLayout<BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float *stage_all = reinterpret_cast<float *>(smem + STAGES * L::kStageBytes);          // 4 warps x 32 x 36 floats
    float *col_part = stage_all + 4 * 32 * 36;                                             // [4 warps][BLOCK_N]
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(col_part + 4 * BLOCK_N);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tmem_full = empty_bar + STAGES;
    uint64_t *tmem_empty = tmem_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
    const int num_tiles = n_tiles * ((M + kBlockM - 1) / kBlockM);

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tm_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tm_b) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(2 * BLOCK_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                    mbar_expect_tx(&full_bar[s], L::kStageBytes);
                    uint8_t *sa = smem + s * L::kStageBytes;
                    tma_load_2d(sa, &tm_a, &full_bar[s], kb * kBlockK, m_blk * kBlockM);
#pragma unroll
                    for (int j = 0; j < BLOCK_N / 32; ++j)
                        tma_load_2d(sa + L::kABytes + j * 4096, &tm_b, &full_bar[s], n_blk * BLOCK_N + j * 32, kb * kBlockK);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_major<BLOCK_N, false, true>();
            uint32_t it = 0, local = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
                const uint32_t acc = local & 1;
                mbar_wait(&tmem_empty[acc], ((local >> 1) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tacc = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&full_bar[s], (it / STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = smem_u32(smem + s * L::kStageBytes);
                    const uint32_t sb = sa + L::kABytes;
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k)
                        umma_tf32(tacc, umma_desc_k_sw128(sa + k * kUmmaK * 4), umma_desc_mn_sw128(sb + k * 1024), idesc,
                                  (kb | k) != 0 ? 1u : 0u);
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&tmem_full[acc]);
            }
        }
    } else {
        const int q = warp & 3;
        float *stage_out = stage_all + (warp - 2) * (32 * 36);
        const int sub = lane & 7, rgrp = lane >> 3;
        uint32_t local = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
            const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
            const uint32_t acc = local & 1;
            const size_t col0 = (size_t)n_blk * BLOCK_N;
            float4 mk[8];
            auto fetch_mask = [&](int c) {                 // the 8 gate vectors this lane needs for a 32-column chunk
#pragma unroll
                for (int it8 = 0; it8 < 8; ++it8) {
                    const int grow = m_blk * kBlockM + q * 32 + it8 * 4 + rgrp;
                    const size_t col = col0 + c * 32 + sub * 4;
                    mk[it8] = (grow < M && col < (size_t)N) ? __ldg(reinterpret_cast<const float4 *>(mask + (size_t)grow * N + col))
                                                            : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            fetch_mask(0);                                 // in flight while the accumulator completes
            mbar_wait(&tmem_full[acc], (local >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + acc * BLOCK_N + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
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
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4 *>(stage_out + lane * 36 + j) =
                        make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                    __uint_as_float(r[j + 3]));
                __syncwarp();
                const size_t col = col0 + c * 32 + sub * 4;
                float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int it8 = 0; it8 < 8; ++it8) {
                    const int rr = it8 * 4 + rgrp;
                    float4 v = *reinterpret_cast<const float4 *>(stage_out + rr * 36 + sub * 4);
                    const int grow = m_blk * kBlockM + q * 32 + rr;
                    if (grow < M && col < (size_t)N) {
                        v.x = mk[it8].x > 0.f ? v.x : 0.f; v.y = mk[it8].y > 0.f ? v.y : 0.f;
                        v.z = mk[it8].z > 0.f ? v.z : 0.f; v.w = mk[it8].w > 0.f ? v.w : 0.f;
                        csum.x += v.x; csum.y += v.y; csum.z += v.z; csum.w += v.w;
                        *reinterpret_cast<float4 *>(C + (size_t)grow * N + col) = v;
                    }
                }
#pragma unroll
                for (int o = 8; o < 32; o <<= 1) {
                    csum.x += __shfl_xor_sync(0xffffffffu, csum.x, o); csum.y += __shfl_xor_sync(0xffffffffu, csum.y, o);
                    csum.z += __shfl_xor_sync(0xffffffffu, csum.z, o); csum.w += __shfl_xor_sync(0xffffffffu, csum.w, o);
                }
                if (rgrp == 0)
                    *reinterpret_cast<float4 *>(col_part + (warp - 2) * BLOCK_N + c * 32 + sub * 4) = csum;
                if (c + 1 < BLOCK_N / 32) fetch_mask(c + 1);
                __syncwarp();
            }
            // the accumulator has been read: hand it back before the column sums are finished
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive_plain(&tmem_empty[acc]);
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int t = threadIdx.x - 64;
            for (int cc = t; cc < BLOCK_N; cc += 128) {
                const float v = col_part[cc] + col_part[BLOCK_N + cc] + col_part[2 * BLOCK_N + cc] + col_part[3 * BLOCK_N + cc];
                if (col0 + cc < (size_t)N) colsum[(size_t)m_blk * N + col0 + cc] = v;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");     // col_part is rewritten by the next tile
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(2 * BLOCK_N) : "memory");
    }
}

// ---- host side ---------------------------------------------------------------------------------------------
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

// 2-D fp32 tensor [rows, cols] row-major, box = [box_rows, 32 cols], 128-byte swizzle, zero fill out of bounds
int make_map(CUtensorMap *map, const float *ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return RLIPV2_DENSE_EDRIVER;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : RLIPV2_DENSE_EDRIVER;
}

template <int BLOCK_N, int STAGES, int ACT>
int launch(const CUtensorMap &ta, const CUtensorMap &tb, const float *bias, const uint8_t *rowmask, float *y, int M,
           int N, int K, cudaStream_t stream) {
    using L = SmemLayout<BLOCK_N>;
    constexpr int smem = STAGES * L::kStageBytes + (2 * STAGES + 1) * 8 + 16 + 1024;
    auto kern = linear_tf32_kernel<BLOCK_N, STAGES, ACT>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    dim3 grid(N / BLOCK_N, (M + kBlockM - 1) / kBlockM, 1);
    kern<<<grid, kThreads, smem, stream>>>(ta, tb, bias, rowmask, y, M, N, K);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

template <int BLOCK_N, int ACT>
int launch_persistent(const CUtensorMap &ta, const CUtensorMap &tb, const float *bias, const uint8_t *rowmask, float *y, int M,
                      int N, int K, cudaStream_t stream) {
    constexpr int STAGES = BLOCK_N == 256 ? 4 : 6;
    using L = SmemLayout<BLOCK_N>;
    constexpr int smem = STAGES * L::kStageBytes + 4 * 32 * 36 * 4 + (2 * STAGES + 4) * 8 + 16 + 1024;
    auto kern = linear_tf32_persistent_kernel<BLOCK_N, STAGES, ACT>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    const long long tiles = (long long)(N / BLOCK_N) * ((M + kBlockM - 1) / kBlockM);
    const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
    kern<<<grid, kThreads, smem, stream>>>(ta, tb, bias, rowmask, y, M, N, K);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

// 2-D fp32 tensor [rows, cols] row-major, box = [box_rows, box_cols <= 32], 128-byte swizzle, zero OOB fill
int make_map_box(CUtensorMap *map, const float *ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols,
                 bool mn_major) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return RLIPV2_DENSE_EDRIVER;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * sizeof(float)};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : RLIPV2_DENSE_EDRIVER;
}

template <int BLOCK_N, bool A_MN, bool B_MN, int EPI>
int launch_gemm(const float *a, const float *b, float *c, const float *mask, float *colsum, int M, int N, int K,
                int splits, cudaStream_t stream) {
    constexpr int STAGES = 3;
    using L = SmemLayout<BLOCK_N>;
    constexpr int smem = STAGES * L::kStageBytes + (2 * STAGES + 1) * 8 + 16 + 1024;
    CUtensorMap ta, tb;
    // K-major operand: stored [rows = M|N, cols = K], box 128|BLOCK_N rows x 32 K;  MN-major: stored [rows = K,
    // cols = M|N], box 32 K rows x 32 columns
    int rc = A_MN ? make_map_box(&ta, a, (uint64_t)K, (uint64_t)M, 32, 32, true) : make_map_box(&ta, a, (uint64_t)M, (uint64_t)K, kBlockM, kBlockK, false);
    if (rc) return rc;
    rc = B_MN ? make_map_box(&tb, b, (uint64_t)K, (uint64_t)N, 32, 32, true) : make_map_box(&tb, b, (uint64_t)N, (uint64_t)K, BLOCK_N, kBlockK, false);
    if (rc) return rc;
    auto kern = gemm_tf32_kernel<BLOCK_N, STAGES, A_MN, B_MN, EPI>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    const int num_kb = (K + kBlockK - 1) / kBlockK;
    if (splits < 1) splits = 1;
    if (splits > num_kb) splits = num_kb;
    const int per = (num_kb + splits - 1) / splits;
    const int gz = (num_kb + per - 1) / per;           // every z slice owns >= 1 k-block
    dim3 grid((N + BLOCK_N - 1) / BLOCK_N, (M + kBlockM - 1) / kBlockM, gz);
    kern<<<grid, kThreads, smem, stream>>>(ta, tb, c, mask, colsum, M, N, num_kb, per);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

inline bool aligned16(const void *p) { return (((uintptr_t)p) & 15) == 0; }

}  // namespace

extern "C" {

int rlipv2_dense_wgrad_tf32(const float *g, const float *x, float *dw, int T, int N, int K, int splits, void *stream)
{
    // dw[N,K] += g[T,N]^T . x[T,K]: logical A = g^T (stored [T, N]: MN-major), B = x^T (stored [T, K]: MN-major),
    // contraction over the T rows, split over `splits` CTAs per output tile, partial tiles reduced into dw
    if (T == 0 || N == 0 || K == 0) return 0;
    if (T < 0 || N < 0 || K < 0 || !g || !x || !dw) return RLIPV2_DENSE_EINVAL;
    if ((N % 4) || (K % 4)) return RLIPV2_DENSE_ESHAPE;
    if (!aligned16(g) || !aligned16(x) || !aligned16(dw)) return RLIPV2_DENSE_EALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    if (K % 256 == 0)
        return launch_gemm<256, true, true, EPI_ATOMIC>(g, x, dw, nullptr, nullptr, N, K, T, splits, s);
    return launch_gemm<128, true, true, EPI_ATOMIC>(g, x, dw, nullptr, nullptr, N, K, T, splits, s);
}

int rlipv2_dense_dgrad_tf32(const float *g, const float *w, float *dx, const float *relu_out, float *colsum,
                            int T, int N, int K, void *stream)
{
    // dx[T,K] = g[T,N] . w[N,K]: A = g (K-major: contraction N contiguous), logical B = w^T [K, N] stored as
    // w [N, K] = [contraction rows, output cols]: MN-major.  relu_out != NULL: dx = (relu_out > 0) ? dx : 0 and
    // colsum_partial[ceil(T/128), K] = per-row-tile column sums of the masked dx (fully written).
    if (T == 0 || K == 0) return 0;
    if (T < 0 || N <= 0 || K < 0 || !g || !w || !dx) return RLIPV2_DENSE_EINVAL;
    if ((N % 4) || (K % 4)) return RLIPV2_DENSE_ESHAPE;
    if (!aligned16(g) || !aligned16(w) || !aligned16(dx) || !aligned16(relu_out) || !aligned16(colsum)) return RLIPV2_DENSE_EALIGN;
    if ((relu_out == nullptr) != (colsum == nullptr)) return RLIPV2_DENSE_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    const int pmin = g_persistent_min_tiles.load(std::memory_order_relaxed);
    if (relu_out && pmin > 0 && g_persistent_dgrad.load(std::memory_order_relaxed) && (K % 256) == 0 && (long long)(K / 128) * ((T + kBlockM - 1) / kBlockM) > pmin) {
        // large gated input gradient (encoder FFN): persistent 128 x 256 tiles.  NB the colsum layout is the same
        // [ceil(T/128), K] as the one-tile-per-CTA kernel's.
        constexpr int STAGES = 4;
        using L = SmemLayout<256>;
        constexpr int smem = STAGES * L::kStageBytes + 4 * 32 * 36 * 4 + 4 * 256 * 4 + (2 * STAGES + 4) * 8 + 16 + 1024;
        CUtensorMap ta, tb;
        int rc = make_map_box(&ta, g, (uint64_t)T, (uint64_t)N, kBlockM, kBlockK, false);
        if (rc) return rc;
        rc = make_map_box(&tb, w, (uint64_t)N, (uint64_t)K, 32, 32, true);
        if (rc) return rc;
        auto kern = dgrad_mask_persistent_kernel<STAGES>;
        static bool configured = false;
        if (!configured) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return (int)e;
            configured = true;
        }
        const long long tiles = (long long)(K / 256) * ((T + kBlockM - 1) / kBlockM);
        const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
        kern<<<grid, kThreads, smem, s>>>(ta, tb, dx, relu_out, colsum, T, K, (N + kBlockK - 1) / kBlockK);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return (int)cudaGetLastError();
    }
    if (relu_out)
        return launch_gemm<128, false, true, EPI_MASK>(g, w, dx, relu_out, colsum, T, K, N, 1, s);
    return launch_gemm<128, false, true, EPI_STORE>(g, w, dx, nullptr, nullptr, T, K, N, 1, s);
}

int rlipv2_dense_linear_splitk_tf32(const float *x, const float *w, const float *bias, float *y, int M, int N, int K,
                                    int splits, void *stream)
{
    // y[M,N] = x[M,K] . w[N,K]^T + bias for small-M / long-K problems (ALIF out projections, label-side in-projections,
    // RobertaLayer FFN-down: 4-64 output tiles on 148 SMs): the K axis is split over `splits` CTAs per output tile, partial
    // tiles reduced into the zero-filled y with red.global.add.v4.f32.  Both operands K-major, as in the plain linear.
    if (M == 0 || N == 0) return 0;
    if (M < 0 || N < 0 || K <= 0 || !x || !w || !y) return RLIPV2_DENSE_EINVAL;
    if ((N % 4) || (K % kBlockK)) return RLIPV2_DENSE_ESHAPE;
    if (!aligned16(x) || !aligned16(w) || !aligned16(y) || !aligned16(bias)) return RLIPV2_DENSE_EALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(y, 0, (size_t)M * N * sizeof(float), s);
    if (e != cudaSuccess) return (int)e;
    return launch_gemm<128, false, false, EPI_ATOMIC_BIAS>(x, w, y, bias, nullptr, M, N, K, splits, s);
}

int rlipv2_dense_linear_tf32_supported(int M, int N, int K)
{
    return (M > 0 && N > 0 && K > 0 && (N % 128) == 0 && (K % kBlockK) == 0) ? 1 : 0;
}

int rlipv2_dense_linear_tf32_rowmask(const float *x, const float *w, const float *bias, const unsigned char *rowmask,
                                     float *y, int M, int N, int K, int act, void *stream)
{
    if (M == 0) return 0;
    if (!rlipv2_dense_linear_tf32_supported(M, N, K)) return RLIPV2_DENSE_ESHAPE;
    if (!x || !w || !y) return RLIPV2_DENSE_EINVAL;
    if (((uintptr_t)x | (uintptr_t)w | (uintptr_t)y | (uintptr_t)bias) & 15) return RLIPV2_DENSE_EALIGN;
    if (act != RLIPV2_DENSE_ACT_NONE && act != RLIPV2_DENSE_ACT_RELU && act != RLIPV2_DENSE_ACT_GELU) return RLIPV2_DENSE_EINVAL;
    // Tile / pipeline choice.  A CTA that has its SM to itself is bound by the latency of its own TMA loads: a 3-stage
    // ring keeps 96 KB in flight (~1 us of L2 latency per 96 KB, i.e. ~8 us for the 768 KB a K = 768 tile streams).
    // Grids of at most one CTA per SM (the ALIF / RobertaLayer / decoder linears: 300-1300 rows) therefore run a
    // 6-stage ring, and - while the grid still fits one CTA per SM - 128x64 tiles, which put twice as many SMs to work.
    const long long ctas128 = (long long)(N / 128) * ((M + kBlockM - 1) / kBlockM);
    const int mode = g_small_mode.load(std::memory_order_relaxed);
    int cfg = 0;
    if (mode >= 1 && ctas128 <= kNumSMs) cfg = 1;
    if (mode >= 2 && 2 * ctas128 <= kNumSMs) cfg = 2;
    const int pmin = g_persistent_min_tiles.load(std::memory_order_relaxed);
    const bool persistent = pmin > 0 && ctas128 > pmin;
    const bool wide = persistent && (N % 256) == 0 && ctas128 >= 8 * kNumSMs;      // 128 x 256 tiles when there are plenty
    CUtensorMap ta, tb;
    int rc = make_map(&ta, x, (uint64_t)M, (uint64_t)K, kBlockM);
    if (rc) return rc;
    rc = make_map(&tb, w, (uint64_t)N, (uint64_t)K, persistent ? (wide ? 256 : 128) : (cfg == 2 ? 64 : 128));
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (persistent) {
#define RLIPV2_DISPATCH_PERSISTENT(BN)                                                                                       \
        switch (act) {                                                                                                       \
            case RLIPV2_DENSE_ACT_NONE: return launch_persistent<BN, RLIPV2_DENSE_ACT_NONE>(ta, tb, bias, rowmask, y, M, N, K, s); \
            case RLIPV2_DENSE_ACT_RELU: return launch_persistent<BN, RLIPV2_DENSE_ACT_RELU>(ta, tb, bias, rowmask, y, M, N, K, s); \
            default: return launch_persistent<BN, RLIPV2_DENSE_ACT_GELU>(ta, tb, bias, rowmask, y, M, N, K, s);               \
        }
        if (wide) { RLIPV2_DISPATCH_PERSISTENT(256) }
        RLIPV2_DISPATCH_PERSISTENT(128)
#undef RLIPV2_DISPATCH_PERSISTENT
    }
#define RLIPV2_DISPATCH_ACT(BN, ST)                                                                              \
    switch (act) {                                                                                               \
        case RLIPV2_DENSE_ACT_NONE: return launch<BN, ST, RLIPV2_DENSE_ACT_NONE>(ta, tb, bias, rowmask, y, M, N, K, s); \
        case RLIPV2_DENSE_ACT_RELU: return launch<BN, ST, RLIPV2_DENSE_ACT_RELU>(ta, tb, bias, rowmask, y, M, N, K, s); \
        default: return launch<BN, ST, RLIPV2_DENSE_ACT_GELU>(ta, tb, bias, rowmask, y, M, N, K, s);               \
    }
    if (cfg == 2) { RLIPV2_DISPATCH_ACT(64, 8) }
    if (cfg == 1) { RLIPV2_DISPATCH_ACT(128, 6) }
    RLIPV2_DISPATCH_ACT(128, 3)
#undef RLIPV2_DISPATCH_ACT
}

void rlipv2_dense_set_small_mode(int mode) { g_small_mode.store(mode < 0 ? 0 : (mode > 2 ? 2 : mode), std::memory_order_relaxed); }

int rlipv2_dense_get_small_mode(void) { return g_small_mode.load(std::memory_order_relaxed); }

void rlipv2_dense_set_persistent_min_tiles(int tiles) { g_persistent_min_tiles.store(tiles < 0 ? 0 : tiles, std::memory_order_relaxed); }

int rlipv2_dense_get_persistent_min_tiles(void) { return g_persistent_min_tiles.load(std::memory_order_relaxed); }

void rlipv2_dense_set_persistent_dgrad(int on) { g_persistent_dgrad.store(on ? 1 : 0, std::memory_order_relaxed); }

int rlipv2_dense_get_persistent_dgrad(void) { return g_persistent_dgrad.load(std::memory_order_relaxed); }

int rlipv2_dense_linear_tf32(const float *x, const float *w, const float *bias, float *y, int M, int N, int K,
                             int act, void *stream)
{
    return rlipv2_dense_linear_tf32_rowmask(x, w, bias, nullptr, y, M, N, K, act, stream);
}

const char *rlipv2_dense_error_string(int code)
{
    switch (code) {
        case 0: return "success";
        case RLIPV2_DENSE_EINVAL: return "rlipv2_dense: invalid argument";
        case RLIPV2_DENSE_ESHAPE: return "rlipv2_dense: shape not supported by the tcgen05 kernel (linear: N % 128, K % 32; grads: N % 4, K % 4)";
        case RLIPV2_DENSE_EALIGN: return "rlipv2_dense: pointers must be 16-byte aligned";
        case RLIPV2_DENSE_EDRIVER: return "rlipv2_dense: cuTensorMapEncodeTiled unavailable or failed";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "rlipv2_dense: unknown error";
    }
}

unsigned long long rlipv2_dense_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
