/*
 * rlipv2_lsap.h - C ABI of the on-device assignment solver of the matcher (SURVEY.md section 8f rank 2,
 * "matcher on device").
 *
 * Replaces, for the graphed train step, the host round trip of
 *     /root/reference/models/matcher.py:185-193
 *         C = C.view(bs, num_queries, -1).cpu()
 *         indices = [linear_sum_assignment(c[i]) for i, c in enumerate(C.split(sizes, -1))]
 * (device->host copy of the cost tensor, scipy's rectangular LSAP per image on the host, host->device copy of the
 * matched indices; once per decoder level, i.e. 3x per step) by one kernel launch that solves every
 * (decoder level, image) problem where the cost tensor already lives.  The index outputs must equal scipy's bit for
 * bit (north_star: "bit-exact for matcher index outputs"): csrc/lsap_core.h restates scipy's shortest-augmenting-path
 * solver (scipy/optimize/rectangular_lsap/rectangular_lsap.cpp - third-party arithmetic)
 * with its tie rules, in double precision, and tests/test_lsap_core.py checks that very header against the installed
 * scipy on the host.
 *
 * Conventions as in rlipv2_msda.h: device pointers, `stream` = cudaStream_t as void*, asynchronous, capturable in a
 * CUDA graph; return 0 / positive cudaError_t / negative RLIPV2_LSAP_E*.  No CPU version.
 */
#ifndef RLIPV2_LSAP_H_
#define RLIPV2_LSAP_H_

#ifdef __cplusplus
extern "C" {
#endif

#define RLIPV2_LSAP_EINVAL (-1)
#define RLIPV2_LSAP_ESIZE (-2)     /* a problem's scratch (8(nr + 2nc) + 4(nr + 3nc) + nr + nc bytes) exceeds 48 KB */

/* cost      [n_levels, bs, nq, T] fp32, dense: the stacked cost tensor of HungarianMatcherHOI (matcher.py:165-185),
 *           T = total number of ground-truth triplets in the batch, image b owning columns
 *           [tgt_start[b], tgt_start[b] + tgt_count[b])
 * tgt_start, tgt_count  [bs] int32 (device)
 * out_offset [n_levels * bs] int64 (device): where problem (level, image) writes its min(nq, tgt_count[b]) pairs
 * out_query, out_target  int64 (device): scipy's (row_ind, col_ind) of every problem - query indices ascending,
 *           target indices relative to the image's first column (what matcher.py:193-194 returns per image)
 * err       [1] int32 (device), must be zeroed by the caller once: set to 1 + problem index if a problem is infeasible or
 *           holds a non-finite cost (scipy raises ValueError there); the outputs of that problem are then undefined
 * max_count upper bound of tgt_count[] known on the host (sizes the scratch; no device->host sync)            */
int rlipv2_lsap_f32(const float *cost, int n_levels, int bs, int nq, int T, const int *tgt_start, const int *tgt_count,
                    int max_count, const long long *out_offset, long long *out_query, long long *out_target, int *err,
                    void *stream);

const char *rlipv2_lsap_error_string(int code);
unsigned long long rlipv2_lsap_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
