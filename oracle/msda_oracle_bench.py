"""TEST INFRASTRUCTURE - timing of the MSDeformAttn CPU port (oracle/msda_oracle.c) on a bounded
sample of the msda_step workload, for bench.py's cpu_baseline / --impl reference legs."""
import time

import numpy as np

from oracle import msda_oracle

LEVELS = [(100, 167), (50, 84), (25, 42), (13, 21)]
M, D, L, P = 8, 32, 4, 4


def time_msda_step_sample(threads, sample_queries=2048, batch=2, num_queries=300):
    S = sum(h * w for h, w in LEVELS)
    rng = np.random.default_rng(0)
    value = rng.standard_normal((1, S, M, D), dtype=np.float32)
    shapes = np.array(LEVELS, dtype=np.int64)
    lsi = np.concatenate(([0], np.cumsum(shapes.prod(1))[:-1])).astype(np.int64)
    Lq = sample_queries
    loc = rng.random((1, Lq, M, L, P, 2), dtype=np.float32)
    attn = rng.random((1, Lq, M, L, P), dtype=np.float32)
    attn /= attn.sum((-1, -2), keepdims=True)
    gout = rng.standard_normal((1, Lq, M * D), dtype=np.float32)
    msda_oracle.forward(value, shapes, lsi, loc[:, :64], attn[:, :64], threads=threads)      # warm
    t0 = time.perf_counter()
    msda_oracle.forward(value, shapes, lsi, loc, attn, threads=threads)
    t1 = time.perf_counter()
    nb = max(1, Lq // max(1, threads))            # backward port parallelises over images only
    msda_oracle.backward(value, shapes, lsi, loc[:, :nb], attn[:, :nb], gout[:, :nb], threads=1)
    t2 = time.perf_counter()
    q_total = batch * (6 * S + 3 * num_queries + 3 * num_queries // 2)
    per_q = (t1 - t0) / Lq + (t2 - t1) / nb / max(1, threads)
    t_step = per_q * q_total
    return {"value": batch / t_step, "unit": "images/s", "cores": threads, "kind": "port",
            "sample": f"1 image, {Lq} encoder-shaped queries fwd on all cores + {nb}-query bwd slice, scaled "
                      "linearly in queries to the 12-call step (oracle/msda_oracle.c)"}
