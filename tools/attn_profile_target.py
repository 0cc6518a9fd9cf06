"""ncu target: the fused attention kernels of one train step, launched eagerly (3 warm-up rounds, then one profiled round).
    ncu --set full --clock-control none --import-source on -k regex:attn_ -s 45 -c 15 -o gpurun_out/attn_prof python tools/attn_profile_target.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rlipv2_b200 import attn_abi  # noqa: E402
from tools.attn_microbench import SHAPES  # noqa: E402

torch.manual_seed(0)
seed = torch.tensor([7], dtype=torch.int64, device="cuda")
for rnd in range(4):                                            # 5 shapes x 3 launches (fwd, dS, 3-problem GEMM) per round
    for name, B, H, Tq, Nk, D, bias, p in SHAPES:
        q = torch.randn(B, Tq, H * D, device="cuda")
        k = torch.randn(B, Nk, H * D, device="cuda")
        v = torch.randn(B, Nk, H * D, device="cuda")
        kb = torch.zeros(B, Nk, device="cuda") if bias else None
        out, stats, su = attn_abi.forward(q, k, v, H, kb, D ** -0.5, p, seed, 3)
        attn_abi.backward(q, k, v, H, kb, out, torch.randn_like(out), stats, D ** -0.5, p, su, 3)
torch.cuda.synchronize()
print("done")
