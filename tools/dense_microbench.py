"""tcgen05 TF32 linear vs cuBLAS TF32 (torch) on the hot-path shapes; CUDA events, rotating inputs > L2."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlipv2_b200 import dense_abi  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = True
SHAPES = [  # (name, M, N, K, act)
    ("enc FFN up (2 img)", 44446, 2048, 256, 1), ("enc FFN down", 44446, 256, 2048, 0),
    ("enc value_proj", 44446, 256, 256, 0), ("enc attn_weights", 44446, 128, 256, 0),
    ("ALIF v_proj", 546, 2048, 256, 0), ("ALIF l_proj", 512, 2048, 768, 0), ("ALIF out_l", 512, 768, 2048, 0),
    ("Roberta qkv", 512, 768, 768, 0), ("Roberta FFN up", 512, 3072, 768, 2), ("Roberta FFN down", 512, 768, 3072, 0),
    ("text tower qkv (256 labels x 5 tok)", 1280, 768, 768, 0), ("text tower FFN up", 1280, 3072, 768, 2),
    ("text tower FFN down", 1280, 768, 3072, 0), ("decoder linear (2 x 300 q)", 600, 256, 256, 0),
    ("decoder FFN up", 600, 2048, 256, 1), ("decoder FFN down", 600, 256, 2048, 0),
]


def timeit(fns, iters):
    """small kernels (a few us) are shorter than the python/ctypes launch path: time them as a CUDA-graph replay of
    `iters` back-to-back launches (what the train step does), big ones eagerly"""
    for f in fns[:3]:
        f()
    torch.cuda.synchronize()
    if GRAPH:
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=st):
                for i in range(iters):
                    fns[i % len(fns)]()
        gr.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            gr.replay()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / (3 * iters) * 1e-3
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fns[i % len(fns)]()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


GRAPH = False
only = sys.argv[1] if len(sys.argv) > 1 else None
for name, M, N, K, act in SHAPES:
    if only and only not in name:
        continue
    bytes_ = 4 * (M * K + N * K + M * N)
    copies = max(2, min(8, int(260e6 / bytes_) + 1))
    xs = [torch.randn(M, K, device="cuda") for _ in range(copies)]
    w = torch.randn(N, K, device="cuda") * K ** -0.5
    b = torch.randn(N, device="cuda")
    ours = [lambda x=x: dense_abi.linear_tf32(x, w, b, act) for x in xs]
    if act == 1:
        ref = [lambda x=x: F.relu(F.linear(x, w, b)) for x in xs]
    elif act == 2:
        ref = [lambda x=x: F.gelu(F.linear(x, w, b)) for x in xs]
    else:
        ref = [lambda x=x: F.linear(x, w, b) for x in xs]
    iters = 50 if M > 10000 else 200
    GRAPH = M <= 10000
    fl = 2.0 * M * N * K
    rec = {"shape": name, "M": M, "N": N, "K": K, "act": act}
    keep = dense_abi.small_mode()
    for mode in ((0, 1, 2) if (N // 128) * ((M + 127) // 128) <= 148 else (keep,)):      # small grids: all tile / ring choices
        dense_abi.set_small_mode(mode)
        rec[f"ours_mode{mode}_us"] = timeit(ours, iters) * 1e6
    dense_abi.set_small_mode(keep)
    sp = dense_abi.splitk_splits(M, N, K) if act == 0 else 1
    if sp > 1:                                      # split-K forward (the default for these shapes since round 2)
        rec["splits"] = sp
        rec["ours_splitk_us"] = timeit([lambda x=x: dense_abi.linear_splitk_tf32(x, w, b, sp) for x in xs], iters) * 1e6
    if (N // 128) * ((M + 127) // 128) > 296:      # large grids: the persistent kernel (double-buffered TMEM accumulator)
        keep_p = dense_abi.persistent_min_tiles()
        dense_abi.set_persistent_min_tiles(296)
        rec["ours_persistent_us"] = timeit(ours, iters) * 1e6
        dense_abi.set_persistent_min_tiles(0)
        t_o = timeit(ours, iters)
        dense_abi.set_persistent_min_tiles(keep_p)
    t_o, t_r = timeit(ours, iters), timeit(ref, iters)
    rec.update({"ours_us": t_o * 1e6, "cublas_us": t_r * 1e6, "ours_TFLOPs": fl / t_o / 1e12,
                "cublas_TFLOPs": fl / t_r / 1e12, "ours_GBs": bytes_ / t_o / 1e9, "default_mode": keep})
    print(json.dumps(rec), flush=True)
