"""Synthetic-input generators for the MSDeformAttn measurements (bench.py, tools/, tests).

Shapes and recipes follow SURVEY.md section 8(d): config 5 uses exactly the recipe of the
reference's models/ops/test.py:32-40; the "encoder" generator reproduces what the ParSeDA encoder
feeds the op at initialisation (reference points = cell centres, models/deformable_transformer.py
:803-815; offsets = the ring pattern of ms_deform_attn.py:66-76 plus noise)."""
import math

import torch

LEVELS_MICRO = [(100, 100), (50, 50), (25, 25), (13, 13)]          # config 5, S = 13294
LEVELS_800x1333 = [(100, 167), (50, 84), (25, 42), (13, 21)]        # R50 @ 3x800x1333, S = 22223


def level_tensors(shapes, device):
    sh = torch.as_tensor(shapes, dtype=torch.long, device=device)
    lsi = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
    return sh, lsi


def msda_bytes(N, S, Lq, M=8, D=32, L=4, P=4, itemsize=4):
    """Algorithmic bytes per call (SURVEY.md section 8(d) / BASELINE.md section 3.4)."""
    fwd = itemsize * N * (S * M * D + 2 * Lq * M * L * P + Lq * M * L * P + Lq * M * D)
    bwd = itemsize * N * (Lq * M * D + 2 * S * M * D + 6 * Lq * M * L * P)
    return fwd, bwd


def random_inputs(N, Lq, shapes, M=8, D=32, P=4, seed=3, device="cuda", dtype=torch.float32):
    """models/ops/test.py:37-40: value = rand*0.01, loc = rand, attn = rand+1e-5 normalised."""
    g = torch.Generator(device=device).manual_seed(seed)
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    value = torch.rand(N, S, M, D, generator=g, device=device, dtype=dtype) * 0.01
    loc = torch.rand(N, Lq, M, L, P, 2, generator=g, device=device, dtype=dtype)
    attn = torch.rand(N, Lq, M, L, P, generator=g, device=device, dtype=dtype) + 1e-5
    attn = attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    gout = torch.randn(N, Lq, M * D, generator=g, device=device, dtype=dtype)
    sh, lsi = level_tensors(shapes, device)
    return value, sh, lsi, loc, attn, gout


def encoder_inputs(N, shapes, M=8, D=32, P=4, seed=3, noise_px=1.0, device="cuda", dtype=torch.float32):
    """Encoder self-attention call: Lq == S, one query per cell in raster order, sampling around
    the query's own position on every level."""
    g = torch.Generator(device=device).manual_seed(seed)
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    value = torch.randn(N, S, M, D, generator=g, device=device, dtype=dtype)
    refs = []
    for (H, W) in shapes:
        ys = (torch.arange(H, device=device, dtype=dtype) + 0.5) / H
        xs = (torch.arange(W, device=device, dtype=dtype) + 0.5) / W
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        refs.append(torch.stack((xx.reshape(-1), yy.reshape(-1)), -1))
    ref = torch.cat(refs, 0)                                           # [S, 2] (x, y)
    thetas = torch.arange(M, device=device, dtype=dtype) * (2.0 * math.pi / M)
    ring = torch.stack([thetas.cos(), thetas.sin()], -1)
    ring = ring / ring.abs().max(-1, keepdim=True)[0]                  # [M, 2]
    offs = ring.view(1, 1, M, 1, 1, 2) * torch.arange(1, P + 1, device=device, dtype=dtype).view(1, 1, 1, 1, P, 1)
    offs = offs + noise_px * torch.randn(N, S, M, L, P, 2, generator=g, device=device, dtype=dtype)
    norm = torch.as_tensor([[w, h] for h, w in shapes], device=device, dtype=dtype)   # (W_l, H_l)
    loc = ref.view(1, S, 1, 1, 1, 2) + offs / norm.view(1, 1, 1, L, 1, 2)
    attn = torch.softmax(torch.randn(N, S, M, L * P, generator=g, device=device, dtype=dtype), -1)
    attn = attn.view(N, S, M, L, P)
    gout = torch.randn(N, S, M * D, generator=g, device=device, dtype=dtype)
    sh, lsi = level_tensors(shapes, device)
    return value, sh, lsi, loc.contiguous(), attn.contiguous(), gout


# ---- synthetic train-step inputs (SURVEY.md section 8d recipe).  This module loads none of the C-ABI libraries, so the CPU
# reference arm of bench.py (oracle/parseda_oracle.py) can build the same batch without touching the product's kernels.
def synthetic_text(n_obj=170, n_verb=85):
    """256 label strings: n_obj object names + 'no objects' + n_verb relation names."""
    objs = [f"object kind {i}" for i in range(n_obj)] + ["no objects"]
    verbs = [f"relation {i} with" for i in range(n_verb)]
    return [(objs, verbs)]


def synthetic_batch(batch, height=800, width=1333, n_obj=170, n_verb=85, triplets=5, seed=0, pin=True):
    """Host-side batch: images [B,3,H,W] fp32 + per-image targets (SURVEY.md section 8d recipe)."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(batch, 3, height, width, generator=g)
    targets = []
    for _ in range(batch):
        def boxes():
            return torch.cat([torch.rand(triplets, 2, generator=g) * 0.4 + 0.3,
                              torch.rand(triplets, 2, generator=g) * 0.2 + 0.1], 1)
        verbs = torch.zeros(triplets, n_verb)
        verbs[torch.arange(triplets), torch.randint(0, n_verb, (triplets,), generator=g)] = 1
        targets.append({"obj_labels": torch.randint(0, n_obj, (triplets,), generator=g),
                        "sub_labels": torch.zeros(triplets, dtype=torch.long), "verb_labels": verbs,
                        "sub_boxes": boxes(), "obj_boxes": boxes()})
    if pin and torch.cuda.is_available():
        images = images.pin_memory()
        targets = [{k: v.pin_memory() for k, v in t.items()} for t in targets]
    return images, targets
