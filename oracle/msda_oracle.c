/*
 * TEST INFRASTRUCTURE - NOT PRODUCT CODE.
 *
 * CPU oracle for multi-scale deformable attention (MSDeformAttn) forward and
 * backward: a plain-C restatement of the reference CUDA kernels
 * /root/reference/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:34-159,238-403
 * (see msda_oracle_impl.h for the line-by-line citations).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path
 * (rlipv2_b200/) never does; it fails loudly when its CUDA library is absent.
 *
 * Pinning: tests/test_oracle_msda.py checks this oracle against
 * tests/golden/msda_*.npz, produced by oracle/gen_golden_msda.py from the
 * reference's own pure-PyTorch op `ms_deform_attn_core_pytorch`
 * (models/ops/functions/ms_deform_attn_func.py:47-65) imported from
 * /root/reference, on the shapes/seeds of the reference's models/ops/test.py.
 *
 * The *_mt entry points split the (batch, query) loop over pthreads so that
 * bench.py can time the port on all host cores.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define REAL float
#define SUFFIX f32
#include "msda_oracle_impl.h"
#undef REAL
#undef SUFFIX

#define REAL double
#define SUFFIX f64
#include "msda_oracle_impl.h"
#undef REAL
#undef SUFFIX

/* ---- multi-threaded fp32 forward (bench baseline only) ------------------ */
typedef struct {
    const float *value; const int64_t *shapes; const int64_t *lsi;
    const float *loc; const float *attn;
    int batch, spatial_size, num_heads, channels, num_levels, num_query, num_point;
    float *out; int b; int q0, q1;
} fwd_job_t;

static void *fwd_worker(void *arg)
{
    fwd_job_t *j = (fwd_job_t *)arg;
    /* one image, a slice of queries: re-base the pointers and call the scalar oracle */
    const int64_t MD = (int64_t)j->num_heads * j->channels;
    const int64_t LP = (int64_t)j->num_levels * j->num_point;
    const int64_t row = ((int64_t)j->b * j->num_query + j->q0) * j->num_heads;
    msda_oracle_forward_f32(j->value + (int64_t)j->b * j->spatial_size * MD, j->shapes, j->lsi,
                            j->loc + row * LP * 2, j->attn + row * LP, 1, j->spatial_size,
                            j->num_heads, j->channels, j->num_levels, j->q1 - j->q0,
                            j->num_point, j->out + row * j->channels);
    return NULL;
}

void msda_oracle_forward_f32_mt(const float *value, const int64_t *spatial_shapes,
                                const int64_t *level_start_index, const float *sampling_loc,
                                const float *attn_weight, int batch, int spatial_size,
                                int num_heads, int channels, int num_levels, int num_query,
                                int num_point, float *out, int num_threads)
{
    if (num_threads < 1) num_threads = 1;
    const int per_img = num_threads;           /* slices per image */
    const int njobs = batch * per_img;
    fwd_job_t *jobs = (fwd_job_t *)calloc((size_t)njobs, sizeof(fwd_job_t));
    pthread_t *th = (pthread_t *)calloc((size_t)num_threads, sizeof(pthread_t));
    int k = 0;
    for (int b = 0; b < batch; ++b)
        for (int s = 0; s < per_img; ++s, ++k) {
            fwd_job_t j = {value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                           batch, spatial_size, num_heads, channels, num_levels, num_query,
                           num_point, out, b,
                           (int)((int64_t)num_query * s / per_img),
                           (int)((int64_t)num_query * (s + 1) / per_img)};
            jobs[k] = j;
        }
    for (int base = 0; base < njobs; base += num_threads) {
        int n = njobs - base < num_threads ? njobs - base : num_threads;
        for (int t = 0; t < n; ++t) pthread_create(&th[t], NULL, fwd_worker, &jobs[base + t]);
        for (int t = 0; t < n; ++t) pthread_join(th[t], NULL);
    }
    free(jobs); free(th);
}

/* ---- multi-threaded fp32 backward: one image per job (grad_value of
 * different images never alias), images round-robin over threads ---------- */
typedef struct {
    const float *value; const int64_t *shapes; const int64_t *lsi;
    const float *loc; const float *attn; const float *gout;
    int spatial_size, num_heads, channels, num_levels, num_query, num_point;
    float *gvalue; float *gloc; float *gattn; int b;
} bwd_job_t;

static void *bwd_worker(void *arg)
{
    bwd_job_t *j = (bwd_job_t *)arg;
    const int64_t MD = (int64_t)j->num_heads * j->channels;
    const int64_t LP = (int64_t)j->num_levels * j->num_point;
    const int64_t row = (int64_t)j->b * j->num_query * j->num_heads;
    msda_oracle_backward_f32(j->value + (int64_t)j->b * j->spatial_size * MD, j->shapes, j->lsi,
                             j->loc + row * LP * 2, j->attn + row * LP,
                             j->gout + row * j->channels, 1, j->spatial_size, j->num_heads,
                             j->channels, j->num_levels, j->num_query, j->num_point,
                             j->gvalue + (int64_t)j->b * j->spatial_size * MD,
                             j->gloc + row * LP * 2, j->gattn + row * LP);
    return NULL;
}

void msda_oracle_backward_f32_mt(const float *value, const int64_t *spatial_shapes,
                                 const int64_t *level_start_index, const float *sampling_loc,
                                 const float *attn_weight, const float *grad_out, int batch,
                                 int spatial_size, int num_heads, int channels, int num_levels,
                                 int num_query, int num_point, float *grad_value,
                                 float *grad_sampling_loc, float *grad_attn_weight,
                                 int num_threads)
{
    if (num_threads < 1) num_threads = 1;
    bwd_job_t *jobs = (bwd_job_t *)calloc((size_t)batch, sizeof(bwd_job_t));
    pthread_t *th = (pthread_t *)calloc((size_t)num_threads, sizeof(pthread_t));
    for (int b = 0; b < batch; ++b) {
        bwd_job_t j = {value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                       grad_out, spatial_size, num_heads, channels, num_levels, num_query,
                       num_point, grad_value, grad_sampling_loc, grad_attn_weight, b};
        jobs[b] = j;
    }
    for (int base = 0; base < batch; base += num_threads) {
        int n = batch - base < num_threads ? batch - base : num_threads;
        for (int t = 0; t < n; ++t) pthread_create(&th[t], NULL, bwd_worker, &jobs[base + t]);
        for (int t = 0; t < n; ++t) pthread_join(th[t], NULL);
    }
    free(jobs); free(th);
}
