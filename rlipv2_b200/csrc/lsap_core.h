// Rectangular linear sum assignment - the arithmetic core shared by the sm_100a kernel (lsap.cu) and the host test
// shim (tests/lsap_host_shim.cpp, compiled with g++ to check this very code against scipy on the build box).
//
// What it must reproduce: `scipy.optimize.linear_sum_assignment`, the call that decides the matcher's index outputs
// (/root/reference/models/matcher.py:16,193; SURVEY.md section 8c(iii): third-party arithmetic, scipy unpinned).  scipy's
// solver (scipy/optimize/rectangular_lsap/rectangular_lsap.cpp, Crouse's shortest-augmenting-path variant of
// Jonker-Volgenant) is restated here from its published algorithm, keeping every rule that decides a tie, because the
// index outputs must be bit-exact:
//   * costs are promoted to double; all dual arithmetic is double add / subtract / compare (no products: nothing for
//     the compiler to contract);
//   * a problem with fewer columns than rows is solved transposed;
//   * the candidate list `remaining` starts as nc-1, nc-2, ..., 0 and a scanned column is removed by moving the last
//     candidate into its slot;
//   * among candidates of equal reduced path cost the scan keeps the first one met unless a later one is unassigned
//     (a sink), in which case the last unassigned one met wins.
// The scan is the only O(nc) inner loop.  It is written per lane (`lane`, `lanes`: lane l visits candidates l, l + lanes,
// ...) and the lanes' partial results are merged with `lsap_beats`, a strict total order that encodes the rule above, so
// the merged winner is the winner of scipy's serial scan whatever the number of lanes or the merge order.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define LSAP_HD __host__ __device__ __forceinline__
#else
#define LSAP_HD inline
#endif

#define LSAP_INF HUGE_VAL

struct LsapBest {
    double val;   // reduced path cost of the lane's best candidate
    int free;     // 1 if that column is unassigned
    int it;       // its slot in `remaining` (-1: the lane saw no candidate)
};

// does a win over b?  lower cost; then an unassigned column over an assigned one; then, between unassigned columns the
// LATER slot, between assigned columns the EARLIER slot (what the serial scan's `<` / `== && unassigned` test leaves)
LSAP_HD bool lsap_beats(const LsapBest &a, const LsapBest &b)
{
    if (a.it < 0) return false;
    if (b.it < 0) return true;
    if (a.val != b.val) return a.val < b.val;
    if (a.free != b.free) return a.free > b.free;
    return a.free ? a.it > b.it : a.it < b.it;
}

struct LsapView {            // one problem: rows x cols entries of a strided fp32 matrix
    const float *base;
    long long row_stride, col_stride;
    int nr, nc;              // the solver's orientation (nr <= nc)
    int transposed;          // 1: solver rows are the matrix's columns
};

LSAP_HD LsapView lsap_view(const float *base, int rows, int cols, long long row_stride, long long col_stride)
{
    LsapView w;
    w.base = base;
    w.transposed = cols < rows;
    if (w.transposed) { w.nr = cols; w.nc = rows; w.row_stride = col_stride; w.col_stride = row_stride; }
    else { w.nr = rows; w.nc = cols; w.row_stride = row_stride; w.col_stride = col_stride; }
    return w;
}

LSAP_HD double lsap_cost(const LsapView &w, int i, int j)
{
    return (double)w.base[(long long)i * w.row_stride + (long long)j * w.col_stride];
}

struct LsapWork {            // caller-provided scratch (shared memory on the device)
    double *u;               // [nr] row duals
    double *v;               // [nc] column duals
    double *sp;              // [nc] shortest path cost to each column in the current search
    int *path;               // [nc] predecessor row
    int *col4row;            // [nr]
    int *row4col;            // [nc]
    int *remaining;          // [nc] unscanned columns
    unsigned char *SR;       // [nr] rows reached
    unsigned char *SC;       // [nc] columns scanned
};

LSAP_HD size_t lsap_work_bytes(int nr, int nc)
{
    size_t d = (size_t)(nr + 2 * nc) * sizeof(double);
    size_t i = (size_t)(nr + 3 * nc) * sizeof(int);
    return d + i + (size_t)(nr + nc);
}

LSAP_HD LsapWork lsap_carve(void *mem, int nr, int nc)
{
    LsapWork W;
    double *d = (double *)mem;
    W.u = d; W.v = d + nr; W.sp = W.v + nc;
    int *p = (int *)(W.sp + nc);
    W.path = p; W.col4row = p + nc; W.row4col = W.col4row + nr; W.remaining = W.row4col + nc;
    W.SR = (unsigned char *)(W.remaining + nc);
    W.SC = W.SR + nr;
    return W;
}

// ---- phases; each `_lane` function is executed by every lane, the others by one ---------------------------------------
LSAP_HD void lsap_init_lane(const LsapWork &W, int nr, int nc, int lane, int lanes)
{
    for (int i = lane; i < nr; i += lanes) { W.u[i] = 0.0; W.col4row[i] = -1; }
    for (int j = lane; j < nc; j += lanes) { W.v[j] = 0.0; W.row4col[j] = -1; }
}

LSAP_HD void lsap_begin_row_lane(const LsapWork &W, int nr, int nc, int lane, int lanes)
{
    for (int i = lane; i < nr; i += lanes) W.SR[i] = 0;
    for (int j = lane; j < nc; j += lanes) { W.SC[j] = 0; W.sp[j] = LSAP_INF; W.remaining[j] = nc - j - 1; }
}

// relax the candidates of this lane from row i and report the lane's best
LSAP_HD LsapBest lsap_scan_lane(const LsapView &w, const LsapWork &W, int i, double min_val, int num_remaining, int lane,
                                int lanes)
{
    LsapBest best;
    best.val = LSAP_INF; best.free = 0; best.it = -1;
    const double ui = W.u[i];
    for (int it = lane; it < num_remaining; it += lanes) {
        const int j = W.remaining[it];
        const double r = min_val + lsap_cost(w, i, j) - ui - W.v[j];
        if (r < W.sp[j]) { W.path[j] = i; W.sp[j] = r; }
        LsapBest c;
        c.val = W.sp[j]; c.free = W.row4col[j] == -1; c.it = it;
        // a lane's own candidates arrive in slot order, so the serial rule applies within the lane as well
        if (lsap_beats(c, best)) best = c;
    }
    return best;
}

// one thread: take the winner out of the candidate list; -> the sink column or -1 (then *i is the next row to expand)
LSAP_HD int lsap_commit(const LsapWork &W, const LsapBest &best, int *i, int *num_remaining)
{
    const int j = W.remaining[best.it];
    int sink = -1;
    if (W.row4col[j] == -1) sink = j;
    else *i = W.row4col[j];
    W.SC[j] = 1;
    W.remaining[best.it] = W.remaining[--(*num_remaining)];
    return sink;
}

LSAP_HD void lsap_update_duals_lane(const LsapWork &W, int nr, int nc, int cur_row, double min_val, int lane, int lanes)
{
    for (int i = lane; i < nr; i += lanes) {
        if (i == cur_row) W.u[i] += min_val;
        else if (W.SR[i]) W.u[i] += min_val - W.sp[W.col4row[i]];
    }
    for (int j = lane; j < nc; j += lanes)
        if (W.SC[j]) W.v[j] -= min_val - W.sp[j];
}

// one thread: flip the assignments along the augmenting path that ends in `sink`
LSAP_HD void lsap_augment(const LsapWork &W, int cur_row, int sink)
{
    int j = sink;
    for (;;) {
        const int i = W.path[j];
        W.row4col[j] = i;
        const int t = W.col4row[i];
        W.col4row[i] = j;
        j = t;
        if (i == cur_row) break;
    }
}

// (row_ind, col_ind) of the caller's matrix, rows ascending, as scipy returns them; min(rows, cols) pairs
LSAP_HD void lsap_emit_lane(const LsapView &w, const LsapWork &W, long long *out_row, long long *out_col, int lane, int lanes)
{
    if (!w.transposed) {
        for (int i = lane; i < w.nr; i += lanes) { out_row[i] = i; out_col[i] = W.col4row[i]; }
        return;
    }
    // solver rows are the matrix's columns: order the pairs by the matrix row (col4row values are distinct)
    for (int t = lane; t < w.nr; t += lanes) {
        const int q = W.col4row[t];
        int rank = 0;
        for (int s = 0; s < w.nr; ++s) rank += W.col4row[s] < q;
        out_row[rank] = q;
        out_col[rank] = t;
    }
}
