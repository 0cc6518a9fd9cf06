"""Tiny driver for ncu: runs `--iters` forward+backward launches of one micro-bench case."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlipv2_b200 import synth  # noqa: E402
from rlipv2_b200.dropin import MultiScaleDeformableAttention as MSDA  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--case", default="enc2")
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--ref", action="store_true")
a = ap.parse_args()
if a.case == "enc2":
    s = synth.encoder_inputs(2, synth.LEVELS_800x1333, seed=0)
elif a.case == "dec16":
    s = synth.random_inputs(16, 300, synth.LEVELS_MICRO, seed=0)
elif a.case == "dec2":
    s = synth.random_inputs(2, 300, synth.LEVELS_MICRO, seed=0)
else:
    s = synth.random_inputs(2, sum(h * w for h, w in synth.LEVELS_MICRO), synth.LEVELS_MICRO, seed=0)
mod = MSDA
if a.ref:
    from oracle import build_ref
    mod = build_ref.load()
for _ in range(a.iters):
    mod.ms_deform_attn_forward(s[0], s[1], s[2], s[3], s[4], 64)
    mod.ms_deform_attn_backward(s[0], s[1], s[2], s[3], s[4], s[5], 64)
torch.cuda.synchronize()
