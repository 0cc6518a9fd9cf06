// TEST INFRASTRUCTURE - host build of rlipv2_b200/csrc/lsap_core.h (g++ -shared), so the arithmetic core of the device
// matcher can be checked against scipy on the build box, which has no GPU.  The driver below runs the phases exactly as
// the kernel in rlipv2_b200/csrc/lsap.cu does, with `lanes` emulated lanes executed one after the other and their scan
// results merged in the kernel's butterfly order.  Not linked into any product library.
#include <stdlib.h>
#include <string.h>

#include "lsap_core.h"

extern "C" int lsap_host_solve(const float *cost, int rows, int cols, long long row_stride, long long col_stride, int lanes,
                               long long *out_row, long long *out_col)
{
    if (rows == 0 || cols == 0) return 0;
    if (lanes < 1 || lanes > 64 || (lanes & (lanes - 1))) return -2;
    const LsapView w = lsap_view(cost, rows, cols, row_stride, col_stride);
    void *mem = malloc(lsap_work_bytes(w.nr, w.nc) + 16);
    const LsapWork W = lsap_carve(mem, w.nr, w.nc);
    int rc = 0;
    for (int l = 0; l < lanes; ++l) lsap_init_lane(W, w.nr, w.nc, l, lanes);
    for (int cur = 0; cur < w.nr && rc == 0; ++cur) {
        for (int l = 0; l < lanes; ++l) lsap_begin_row_lane(W, w.nr, w.nc, l, lanes);
        int i = cur, sink = -1, num_remaining = w.nc;
        double min_val = 0.0;
        while (sink == -1) {
            W.SR[i] = 1;
            LsapBest b[64];
            for (int l = 0; l < lanes; ++l) b[l] = lsap_scan_lane(w, W, i, min_val, num_remaining, l, lanes);
            for (int off = lanes / 2; off > 0; off >>= 1)          // __shfl_xor butterfly: every lane ends with the winner
                for (int l = 0; l < lanes; ++l) {
                    const LsapBest o = b[l ^ off];
                    if (l < (l ^ off)) {                           // process each pair once, symmetric result
                        const LsapBest win = lsap_beats(o, b[l]) ? o : b[l];
                        b[l] = win; b[l ^ off] = win;
                    }
                }
            min_val = b[0].val;
            if (b[0].it < 0 || !(min_val < LSAP_INF)) { rc = -1; break; }   // infeasible / non-finite costs
            sink = lsap_commit(W, b[0], &i, &num_remaining);
        }
        if (rc) break;
        for (int l = 0; l < lanes; ++l) lsap_update_duals_lane(W, w.nr, w.nc, cur, min_val, l, lanes);
        lsap_augment(W, cur, sink);
    }
    if (rc == 0)
        for (int l = 0; l < lanes; ++l) lsap_emit_lane(w, W, out_row, out_col, l, lanes);
    free(mem);
    return rc;
}
