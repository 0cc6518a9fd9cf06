"""How many grad_value reductions of the fast MSDeformAttn backward the within-pair merge removes (CPU, no GPU needed).

For the encoder-shaped inputs of `rlipv2_b200.synth.encoder_inputs` (one query per cell, ring offsets + `noise` cells of Gaussian
noise) it counts, per (image, query, head) pair: valid corners (= reductions of the unmerged schedule), distinct cells per level
(= reductions of the merged schedule, rlipv2_b200/csrc/msda_merge.h) and weightless corners (exact-integer sample positions).
The numbers quoted in profiles/msda_r02.md come from here:

    python tools/msda_merge_sim.py            # noise 1.0, 0.0, 0.25
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlipv2_b200 import synth  # noqa: E402


def simulate(noise, shapes=synth.LEVELS_800x1333):
    _, _, _, loc, _, _ = synth.encoder_inputs(1, shapes, noise_px=noise, device="cpu")
    _, S, M, L, P, _ = loc.shape
    valid_n = unique_n = zero_n = 0
    for lvl, (H, W) in enumerate(shapes):
        x = loc[0, :, :, lvl, :, 0] * W - 0.5
        y = loc[0, :, :, lvl, :, 1] * H - 0.5
        inside = (x > -1) & (y > -1) & (x < W) & (y < H)
        x0, y0 = torch.floor(x).long(), torch.floor(y).long()
        lw, lh = x - x0, y - y0
        cells, valid, zero = [], [], []
        for cy in (0, 1):
            for cx in (0, 1):
                yy, xx = y0 + cy, x0 + cx
                v = inside & (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
                w = (lh if cy else 1 - lh) * (lw if cx else 1 - lw)
                cells.append(yy * 100000 + xx)
                valid.append(v)
                zero.append(v & (w == 0))
        cells = torch.stack(cells, -1).reshape(S, M, P * 4)
        valid = torch.stack(valid, -1).reshape(S, M, P * 4)
        zero = torch.stack(zero, -1).reshape(S, M, P * 4)
        cells = torch.where(valid, cells, torch.full_like(cells, -1))
        srt, _ = cells.sort(-1)
        uniq = ((srt[..., 1:] != srt[..., :-1]) & (srt[..., 1:] >= 0)).sum(-1) + (srt[..., 0] >= 0).long()
        valid_n += int(valid.sum())
        unique_n += int(uniq.sum())
        zero_n += int(zero.sum())
    pairs = S * M
    return valid_n / pairs, unique_n / pairs, zero_n / pairs


if __name__ == "__main__":
    for noise in (1.0, 0.0, 0.25):
        v, u, z = simulate(noise)
        print(f"noise {noise:4.2f} cell(s): valid corners / pair {v:5.1f}, distinct cells / pair {u:5.1f} ({100 * (1 - u / v):4.1f} % fewer "
              f"reductions), weightless corners / pair {z:4.1f}")
