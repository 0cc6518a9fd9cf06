"""Micro-benchmark of the fused tcgen05 attention cores against the torch / cuBLAS TF32 op chain they replace
(bmm -> softmax -> dropout -> bmm), on the attention shapes of one RLIPv2-ParSeDA R50 train step (BASELINE config 2, batch 2).
Timed as CUDA-graph replays (the kernels are shorter than the python launch path), CUDA events, forward and forward+backward.
    python tools/attn_microbench.py > profiles/attn_microbench_r02.jsonl"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

SHAPES = [  # name, B, H, Tq, Nk, D, key bias, dropout
    ("alif_v2l", 2, 8, 273, 256, 256, False, 0.1),
    ("alif_l2v", 2, 8, 256, 273, 256, False, 0.1),
    ("roberta", 2, 12, 256, 256, 64, True, 0.1),
    ("pair_dec", 2, 8, 300, 300, 32, False, 0.0),
    ("verb_dec", 2, 8, 150, 150, 32, False, 0.0),
]


def timed_graph(fn, reps=10, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps):
                fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * iters) * 1e3          # us per call


def main():
    from rlipv2_b200 import dense
    torch.manual_seed(0)
    for name, B, H, Tq, Nk, D, bias, p in SHAPES:
        C = H * D
        q = torch.randn(B, Tq, C, device="cuda", requires_grad=True)
        k = torch.randn(B, Nk, C, device="cuda", requires_grad=True)
        v = torch.randn(B, Nk, C, device="cuda", requires_grad=True)
        go = torch.randn(B, Tq, C, device="cuda")
        kb = None
        if bias:
            kb = torch.zeros(B, Nk, device="cuda")
            kb[:, -20:] = -10000.0
        row = {"shape": name, "B": B, "H": H, "Tq": Tq, "Nk": Nk, "D": D, "dropout": p}
        flops_f = 4.0 * B * H * Tq * Nk * D
        for mode, tag in (("tf32", "fused"), ("fp32", None)):
            dense.set_matmul_precision(mode)
            if tag is None:                                   # torch path with TF32 products allowed = cuBLAS TF32
                torch.backends.cuda.matmul.allow_tf32 = True
                tag = "torch_tf32"

            def fwd():
                with torch.no_grad():
                    return dense.attention(q, k, v, H, D ** -0.5, kb, p, True, 3)

            def fwd_bwd():
                o = dense.attention(q, k, v, H, D ** -0.5, kb, p, True, 3)
                gq, gk, gv = torch.autograd.grad(o, (q, k, v), go)
                return gq

            row[tag + "_fwd_us"] = round(timed_graph(fwd), 2)
            row[tag + "_fwd_bwd_us"] = round(timed_graph(fwd_bwd), 2)
        dense.set_matmul_precision("fp32")
        row["fused_fwd_tflops"] = round(flops_f / row["fused_fwd_us"] * 1e-6, 2)
        row["fused_fwd_bwd_tflops"] = round(3.5 * flops_f / row["fused_fwd_bwd_us"] * 1e-6, 2)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
