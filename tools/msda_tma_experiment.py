"""Default MSDeformAttn forward vs the TMA-staged coarse-level variant on the encoder call of a 800x1333 image (N = 2,
S = Lq = 22223); CUDA events, three rotating input sets (> L2).   python tools/msda_tma_experiment.py [--once]
(--once: one launch of each after warm-up, for ncu)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rlipv2_b200 import msda_abi, synth  # noqa: E402
from rlipv2_b200.dropin import MultiScaleDeformableAttention as MSDA  # noqa: E402

shapes = synth.LEVELS_800x1333
coarse = sum(h * w for h, w in shapes[:2])
sets = [synth.encoder_inputs(2, shapes, seed=s) for s in range(3)]
fns = {"default": lambda s: MSDA.ms_deform_attn_forward(s[0], s[1], s[2], s[3], s[4], 64),
       "tma_coarse_levels": lambda s: msda_abi.forward_tma(s[0], s[1], s[2], s[3], s[4], coarse)}
if "--once" in sys.argv:
    for f in fns.values():
        for s in sets:
            f(s)
    torch.cuda.synchronize()
    sys.exit(0)
fwd_b, _ = synth.msda_bytes(2, sum(h * w for h, w in shapes), sum(h * w for h, w in shapes))
for name, f in fns.items():
    for i in range(5):
        f(sets[i % 3])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(30):
        f(sets[i % 3])
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / 30 * 1e3
    print(json.dumps({"kernel": name, "us": round(us, 1), "algorithmic_GBs": round(fwd_b / us * 1e-3, 1)}), flush=True)
