// lsap.cu - the matcher's assignment problems solved on the device (sm_100a).  See include/rlipv2_lsap.h for what it
// replaces and csrc/lsap_core.h for the algorithm (shared with the host test shim).
//
// One warp per (decoder level, image) problem: 150 queries x a handful of ground-truth triplets, i.e. a few hundred
// dependent steps of double-precision dual updates - latency-bound work for which one resident warp with its state in
// shared memory is the right size.  The only O(columns) loop, the relaxation scan, is split across the 32 lanes and
// merged by a shuffle butterfly under `lsap_beats`, a total order that reproduces the serial scan's tie rules; the
// bookkeeping between scans is done by lane 0.  All problems of a step run side by side in one launch.
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "lsap_core.h"
#include "rlipv2_lsap.h"

namespace {

std::atomic<unsigned long long> g_launches{0};
constexpr size_t kMaxScratch = 48 * 1024;

__device__ __forceinline__ LsapBest warp_best(LsapBest b)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        LsapBest o;
        o.val = __shfl_xor_sync(0xffffffffu, b.val, off);
        o.free = __shfl_xor_sync(0xffffffffu, b.free, off);
        o.it = __shfl_xor_sync(0xffffffffu, b.it, off);
        if (lsap_beats(o, b)) b = o;
    }
    return b;                                         // identical in every lane: the order is total
}

__global__ void __launch_bounds__(32)
lsap_kernel(const float *__restrict__ cost, int bs, int nq, int T, const int *__restrict__ tgt_start,
            const int *__restrict__ tgt_count, const long long *__restrict__ out_offset, long long *__restrict__ out_query,
            long long *__restrict__ out_target, int *__restrict__ err, unsigned scratch_bytes)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x;
    const int p = blockIdx.x;                         // problem = level * bs + image
    const int b = p % bs;
    const int n = tgt_count[b];
    if (n <= 0 || nq <= 0) return;
    const LsapView w = lsap_view(cost + (size_t)p * nq * T + tgt_start[b], nq, n, T, 1);
    if (lsap_work_bytes(w.nr, w.nc) > scratch_bytes) {              // max_count on the host was not an upper bound
        if (lane == 0) *err = p + 1;
        return;
    }
    const LsapWork W = lsap_carve(smem, w.nr, w.nc);
    // non-finite costs: scipy refuses them; report instead of looping on NaN comparisons
    int bad = 0;
    for (int e = lane; e < w.nr * w.nc; e += 32) {
        const float c = w.base[(long long)(e / w.nc) * w.row_stride + (long long)(e % w.nc) * w.col_stride];
        bad |= !(fabsf(c) <= 3.402823466e38f);
    }
    if (__any_sync(0xffffffffu, bad)) {
        if (lane == 0) *err = p + 1;
        return;
    }
    lsap_init_lane(W, w.nr, w.nc, lane, 32);
    __syncwarp();
    for (int cur = 0; cur < w.nr; ++cur) {
        lsap_begin_row_lane(W, w.nr, w.nc, lane, 32);
        __syncwarp();
        int i = cur, sink = -1, num_remaining = w.nc;
        double min_val = 0.0;
        while (sink == -1) {
            if (lane == 0) W.SR[i] = 1;
            const LsapBest best = warp_best(lsap_scan_lane(w, W, i, min_val, num_remaining, lane, 32));
            min_val = best.val;
            if (best.it < 0 || !(min_val < LSAP_INF)) {             // infeasible (cannot happen with finite costs)
                if (lane == 0) *err = p + 1;
                return;
            }
            __syncwarp();                                           // every lane's relaxations are visible to lane 0
            if (lane == 0) sink = lsap_commit(W, best, &i, &num_remaining);
            sink = __shfl_sync(0xffffffffu, sink, 0);
            i = __shfl_sync(0xffffffffu, i, 0);
            num_remaining = __shfl_sync(0xffffffffu, num_remaining, 0);
            __syncwarp();                                           // and lane 0's list update to every lane
        }
        lsap_update_duals_lane(W, w.nr, w.nc, cur, min_val, lane, 32);
        __syncwarp();
        if (lane == 0) lsap_augment(W, cur, sink);
        __syncwarp();
    }
    lsap_emit_lane(w, W, out_query + out_offset[p], out_target + out_offset[p], lane, 32);
}

}  // namespace

extern "C" {

int rlipv2_lsap_f32(const float *cost, int n_levels, int bs, int nq, int T, const int *tgt_start, const int *tgt_count,
                    int max_count, const long long *out_offset, long long *out_query, long long *out_target, int *err,
                    void *stream)
{
    if (n_levels < 0 || bs < 0 || nq < 0 || T < 0 || max_count < 0) return RLIPV2_LSAP_EINVAL;
    if (n_levels == 0 || bs == 0 || nq == 0 || T == 0 || max_count == 0) return 0;
    if (!cost || !tgt_start || !tgt_count || !out_offset || !out_query || !out_target || !err) return RLIPV2_LSAP_EINVAL;
    const int nr = nq < max_count ? nq : max_count, nc = nq < max_count ? max_count : nq;
    const size_t scratch = (lsap_work_bytes(nr, nc) + 15) & ~(size_t)15;
    if (scratch > kMaxScratch) return RLIPV2_LSAP_ESIZE;
    lsap_kernel<<<n_levels * bs, 32, scratch, (cudaStream_t)stream>>>(cost, bs, nq, T, tgt_start, tgt_count, out_offset,
                                                                      out_query, out_target, err, (unsigned)scratch);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

const char *rlipv2_lsap_error_string(int code)
{
    if (code == RLIPV2_LSAP_EINVAL) return "invalid argument";
    if (code == RLIPV2_LSAP_ESIZE) return "problem too large for the shared-memory scratch";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "ok";
}

unsigned long long rlipv2_lsap_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
