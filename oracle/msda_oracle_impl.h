/*
 * TEST INFRASTRUCTURE - NOT PRODUCT CODE.
 *
 * Type-generic body of the MSDeformAttn CPU oracle.  Included twice by
 * msda_oracle.c with REAL = float / double and SUFFIX = f32 / f64.
 *
 * It restates, loop for loop, the arithmetic of the reference CUDA kernels
 * (all citations relative to /root/reference/models/ops/src/cuda/):
 *   forward : ms_deform_im2col_cuda.cuh:238-299 (kernel) and :34-84 (bilinear)
 *   backward: ms_deform_im2col_cuda.cuh:302-403 (kernel, D=32 dispatch) and
 *             :88-159 (bilinear + scatter)
 * with the one deliberate difference that the scatter into grad_value is a
 * sequential "+=" in (b,q,m,c,l,p) order instead of the reference's fp32
 * atomicAdd (whose order is nondeterministic).
 */

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

/* cuh:34-84.  `data` points at value[b, level_start, 0, 0]; the element for
 * cell (h, w), head m, channel c sits at h*W*M*D + w*M*D + m*D + c (cuh:47-53). */
static REAL FN(bilinear)(const REAL *data, int height, int width, int nheads,
                         int channels, REAL h, REAL w, int m, int c)
{
    const int h_low = (int)floor((double)h);
    const int w_low = (int)floor((double)w);
    const int h_high = h_low + 1;
    const int w_high = w_low + 1;

    const REAL lh = h - (REAL)h_low;
    const REAL lw = w - (REAL)w_low;
    const REAL hh = (REAL)1 - lh, hw = (REAL)1 - lw;

    const int64_t w_stride = (int64_t)nheads * channels;
    const int64_t h_stride = (int64_t)width * w_stride;
    const int64_t base = (int64_t)m * channels + c;

    REAL v1 = 0, v2 = 0, v3 = 0, v4 = 0;
    if (h_low >= 0 && w_low >= 0)
        v1 = data[h_low * h_stride + w_low * w_stride + base];
    if (h_low >= 0 && w_high <= width - 1)
        v2 = data[h_low * h_stride + w_high * w_stride + base];
    if (h_high <= height - 1 && w_low >= 0)
        v3 = data[h_high * h_stride + w_low * w_stride + base];
    if (h_high <= height - 1 && w_high <= width - 1)
        v4 = data[h_high * h_stride + w_high * w_stride + base];

    const REAL w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
    return (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);
}

/* cuh:238-299: out[b,q,m,c] = sum_{l,p} attn * bilinear(value_l, loc*size - 0.5) */
void FN(msda_oracle_forward)(const REAL *value, const int64_t *spatial_shapes,
                             const int64_t *level_start_index, const REAL *sampling_loc,
                             const REAL *attn_weight, int batch, int spatial_size,
                             int num_heads, int channels, int num_levels, int num_query,
                             int num_point, REAL *out)
{
    const int64_t qid_stride = (int64_t)num_heads * channels;
    for (int b = 0; b < batch; ++b)
    for (int q = 0; q < num_query; ++q)
    for (int m = 0; m < num_heads; ++m) {
        const int64_t sampling_index = ((int64_t)b * num_query + q) * num_heads + m;
        for (int c = 0; c < channels; ++c) {
            int64_t wptr = sampling_index * num_levels * num_point;
            int64_t lptr = wptr << 1;
            REAL col = 0;
            for (int l = 0; l < num_levels; ++l) {
                const int64_t level_start = level_start_index[l];
                const int H = (int)spatial_shapes[2 * l];
                const int W = (int)spatial_shapes[2 * l + 1];
                const REAL *data = value + ((int64_t)b * spatial_size + level_start) * qid_stride;
                for (int p = 0; p < num_point; ++p) {
                    const REAL loc_w = sampling_loc[lptr];
                    const REAL loc_h = sampling_loc[lptr + 1];
                    const REAL weight = attn_weight[wptr];
                    const REAL h_im = loc_h * (REAL)H - (REAL)0.5;   /* cuh:285 */
                    const REAL w_im = loc_w * (REAL)W - (REAL)0.5;   /* cuh:286 */
                    if (h_im > -1 && w_im > -1 && h_im < H && w_im < W)   /* cuh:288 */
                        col += FN(bilinear)(data, H, W, num_heads, channels, h_im, w_im, m, c) * weight;
                    wptr += 1;
                    lptr += 2;
                }
            }
            out[sampling_index * channels + c] = col;
        }
    }
}

/* cuh:302-403 with cuh:88-159 inlined.  grad_* must be zero-filled by the
 * caller, as the reference host wrapper does (ms_deform_attn_cuda.cu:121-123). */
void FN(msda_oracle_backward)(const REAL *value, const int64_t *spatial_shapes,
                              const int64_t *level_start_index, const REAL *sampling_loc,
                              const REAL *attn_weight, const REAL *grad_out, int batch,
                              int spatial_size, int num_heads, int channels, int num_levels,
                              int num_query, int num_point, REAL *grad_value,
                              REAL *grad_sampling_loc, REAL *grad_attn_weight)
{
    const int64_t qid_stride = (int64_t)num_heads * channels;
    for (int b = 0; b < batch; ++b)
    for (int q = 0; q < num_query; ++q)
    for (int m = 0; m < num_heads; ++m) {
        const int64_t sampling_index = ((int64_t)b * num_query + q) * num_heads + m;
        for (int c = 0; c < channels; ++c) {
            const REAL top_grad = grad_out[sampling_index * channels + c];
            int64_t wptr = sampling_index * num_levels * num_point;
            int64_t lptr = wptr << 1;
            for (int l = 0; l < num_levels; ++l) {
                const int64_t level_start = level_start_index[l];
                const int height = (int)spatial_shapes[2 * l];
                const int width = (int)spatial_shapes[2 * l + 1];
                const int64_t off = ((int64_t)b * spatial_size + level_start) * qid_stride;
                const REAL *data = value + off;
                REAL *gdata = grad_value + off;
                for (int p = 0; p < num_point; ++p, wptr += 1, lptr += 2) {
                    const REAL loc_w = sampling_loc[lptr];
                    const REAL loc_h = sampling_loc[lptr + 1];
                    const REAL weight = attn_weight[wptr];
                    const REAL h = loc_h * (REAL)height - (REAL)0.5;
                    const REAL w = loc_w * (REAL)width - (REAL)0.5;
                    if (!(h > -1 && w > -1 && h < height && w < width))
                        continue;
                    /* cuh:98-158 */
                    const int h_low = (int)floor((double)h);
                    const int w_low = (int)floor((double)w);
                    const int h_high = h_low + 1;
                    const int w_high = w_low + 1;
                    const REAL lh = h - (REAL)h_low;
                    const REAL lw = w - (REAL)w_low;
                    const REAL hh = (REAL)1 - lh, hw = (REAL)1 - lw;
                    const int64_t w_stride = qid_stride;
                    const int64_t h_stride = (int64_t)width * w_stride;
                    const int64_t base = (int64_t)m * channels + c;
                    const REAL w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
                    const REAL top_grad_value = top_grad * weight;
                    REAL grad_h_weight = 0, grad_w_weight = 0;
                    REAL v1 = 0, v2 = 0, v3 = 0, v4 = 0;
                    if (h_low >= 0 && w_low >= 0) {
                        const int64_t ptr = h_low * h_stride + w_low * w_stride + base;
                        v1 = data[ptr];
                        grad_h_weight -= hw * v1;
                        grad_w_weight -= hh * v1;
                        gdata[ptr] += w1 * top_grad_value;
                    }
                    if (h_low >= 0 && w_high <= width - 1) {
                        const int64_t ptr = h_low * h_stride + w_high * w_stride + base;
                        v2 = data[ptr];
                        grad_h_weight -= lw * v2;
                        grad_w_weight += hh * v2;
                        gdata[ptr] += w2 * top_grad_value;
                    }
                    if (h_high <= height - 1 && w_low >= 0) {
                        const int64_t ptr = h_high * h_stride + w_low * w_stride + base;
                        v3 = data[ptr];
                        grad_h_weight += hw * v3;
                        grad_w_weight -= lh * v3;
                        gdata[ptr] += w3 * top_grad_value;
                    }
                    if (h_high <= height - 1 && w_high <= width - 1) {
                        const int64_t ptr = h_high * h_stride + w_high * w_stride + base;
                        v4 = data[ptr];
                        grad_h_weight += lw * v4;
                        grad_w_weight += lh * v4;
                        gdata[ptr] += w4 * top_grad_value;
                    }
                    const REAL val = (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);
                    /* the per-channel partials the reference parks in shared memory
                     * and sums over the block (cuh:377-393) */
                    grad_attn_weight[wptr] += top_grad * val;
                    grad_sampling_loc[lptr] += (REAL)width * grad_w_weight * top_grad_value;
                    grad_sampling_loc[lptr + 1] += (REAL)height * grad_h_weight * top_grad_value;
                }
            }
        }
    }
}

#undef FN
#undef CAT
#undef CAT_
