"""tcgen05 wgrad / dgrad kernels vs cuBLAS (TF32) on the encoder's backward shapes.  One JSON line per case."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlipv2_b200 import dense_abi, fused_abi  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = True
T = 44446


def t(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(64 << 20, device="cuda")
    tot = 0.0
    for _ in range(iters):
        flush.zero_()                      # 256 MB > L2
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters * 1e3


for name, N, K in [("msda_proj 256x256", 256, 256), ("attn_w 128x256", 128, 256), ("ffn_up dW1 2048x256", 2048, 256),
                   ("ffn_down dW2 256x2048", 256, 2048)]:
    g = torch.randn(T, N, device="cuda")
    x = torch.randn(T, K, device="cuda")
    w = torch.randn(N, K, device="cuda")
    ours = t(lambda: dense_abi.wgrad_tf32(g, x))
    ref = t(lambda: g.t() @ x)
    line = {"case": "wgrad " + name, "ours_us": ours, "cublas_us": ref, "splits": dense_abi.wgrad_splits(T, N, K),
            "hbm_floor_us": 4.0 * T * (N + K) / 6.45e12 * 1e6}
    for sp in (8, 16, 37, 74, 148):
        line[f"ours_us_splits{sp}"] = t(lambda: dense_abi.wgrad_tf32(g, x, sp), 10)
    print(json.dumps(line), flush=True)
    ours = t(lambda: dense_abi.dgrad_tf32(g, w))
    ref = t(lambda: g @ w)
    print(json.dumps({"case": "dgrad " + name, "ours_us": ours, "cublas_us": ref,
                      "hbm_floor_us": 4.0 * T * (N + K) / 6.45e12 * 1e6}), flush=True)
g = torch.randn(T, 256, device="cuda")
w2 = torch.randn(256, 2048, device="cuda")
h = torch.relu(torch.randn(T, 2048, device="cuda"))
ours = t(lambda: dense_abi.dgrad_tf32(g, w2, relu_out=h))
ref = t(lambda: fused_abi.relu_bwd_colsum(g @ w2, h))
print(json.dumps({"case": "dgrad ffn_down + relu mask + colsum", "ours_us": ours, "cublas_plus_mask_us": ref}), flush=True)
