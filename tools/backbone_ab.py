"""A/B: backbone forward+backward time with and without frozen-BN folding (CUDA graphs, so launch overhead is out)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlipv2_b200 import dense, models  # noqa: E402
from rlipv2_b200.nested import NestedTensor  # noqa: E402

dense.set_matmul_precision("tf32")
args = models.default_args(device="cuda", num_queries=300, synthetic_text_encoder=True)
from rlipv2_b200.backbone import build_backbone  # noqa: E402
bb = build_backbone(args).cuda().train()
x = torch.randn(2, 3, 800, 1333, device="cuda")
mask = torch.zeros(2, 800, 1333, dtype=torch.bool, device="cuda")


def run():
    feats, pos = bb(NestedTensor(x, mask))
    loss = sum(f.tensors.square().mean() for f in feats)
    loss.backward()


for fold in (True, False, True, False):
    bb[0].fold_bn = fold
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            run()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    for p in bb.parameters():
        p.grad = None
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        run()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    print(f"fold_bn={fold}: backbone fwd+bwd {a.elapsed_time(b) / 20:.3f} ms")
