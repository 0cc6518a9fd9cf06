"""TEST INFRASTRUCTURE - generator of tests/golden/msda_*.npz.

Runs HERE (the build container, where /root/reference exists), never on the GPU box.
It imports the reference's own pure-PyTorch implementation of the op,
``ms_deform_attn_core_pytorch`` (/root/reference/models/ops/functions/ms_deform_attn_func.py:47-65),
evaluates forward in fp64 and the three input gradients by autograd (through
``F.grid_sample``), and stores inputs + outputs as small .npz fixtures.  The cases follow the
reference's only test, models/ops/test.py: seed 3, ``value = rand*0.01``, ``loc = rand``,
``attn = rand + 1e-5`` normalised over (L, P) (test.py:32-40), channel counts from its gradcheck
list (test.py:88), plus edge cases the kernels guard against (samples outside the map,
cuh:288 and cuh:56-78; ragged non-square levels; a single 1x1 level).

    python oracle/gen_golden_msda.py          # rewrites tests/golden/msda_*.npz
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference_op():
    # the reference module does `import MultiScaleDeformableAttention as MSDA` at import time
    # (ms_deform_attn_func.py:22); the pure-PyTorch function we want never touches it.
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    sys.path.insert(0, os.path.join(REF, "models", "ops"))
    from functions.ms_deform_attn_func import ms_deform_attn_core_pytorch
    return ms_deform_attn_core_pytorch


CASES = {
    # name: (N, M, D, Lq, L, P, shapes, loc_range)
    "reftest_d2": (1, 2, 2, 2, 2, 2, [(6, 4), (3, 2)], (0.0, 1.0)),          # test.py:25-29 verbatim
    "reftest_d30": (1, 2, 30, 2, 2, 2, [(6, 4), (3, 2)], (0.0, 1.0)),        # test.py:88 channel list
    "reftest_d32": (1, 2, 32, 2, 2, 2, [(6, 4), (3, 2)], (0.0, 1.0)),
    "reftest_d64": (1, 2, 64, 2, 2, 2, [(6, 4), (3, 2)], (0.0, 1.0)),
    "reftest_d71": (1, 2, 71, 2, 2, 2, [(6, 4), (3, 2)], (0.0, 1.0)),
    "parseda_small": (2, 8, 32, 19, 4, 4, [(8, 11), (4, 6), (2, 3), (1, 2)], (0.0, 1.0)),   # ParSeDA head layout
    "out_of_range": (2, 8, 32, 11, 4, 4, [(7, 5), (4, 3), (2, 2), (1, 1)], (-0.3, 1.3)),     # cuh:288 guards
    "one_cell": (1, 1, 32, 5, 1, 1, [(1, 1)], (-0.5, 1.5)),
    "lp_odd": (2, 4, 16, 7, 3, 5, [(5, 8), (3, 4), (2, 2)], (-0.1, 1.1)),                    # L*P = 15, D=16
}


def make_inputs(N, M, D, Lq, L, P, shapes, loc_range, seed=3):
    g = torch.Generator().manual_seed(seed)
    S = sum(h * w for h, w in shapes)
    value = torch.rand(N, S, M, D, generator=g, dtype=torch.float64) * 0.01
    lo, hi = loc_range
    loc = torch.rand(N, Lq, M, L, P, 2, generator=g, dtype=torch.float64) * (hi - lo) + lo
    attn = torch.rand(N, Lq, M, L, P, generator=g, dtype=torch.float64) + 1e-5
    attn = attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    grad_out = torch.randn(N, Lq, M * D, generator=g, dtype=torch.float64)
    # fp32-representable inputs, so the fp32 oracle / CUDA kernels see bit-identical operands
    value, loc, attn, grad_out = (t.float().double() for t in (value, loc, attn, grad_out))
    shapes_t = torch.as_tensor(shapes, dtype=torch.long)
    lsi = torch.cat((shapes_t.new_zeros((1,)), shapes_t.prod(1).cumsum(0)[:-1]))
    return value, shapes_t, lsi, loc, attn, grad_out


def main():
    core = import_reference_op()
    os.makedirs(OUT, exist_ok=True)
    for name, cfg in CASES.items():
        value, shapes, lsi, loc, attn, grad_out = make_inputs(*cfg)
        v, l_, a = (t.clone().requires_grad_(True) for t in (value, loc, attn))
        out = core(v, shapes, l_, a)
        out.backward(grad_out)
        np.savez_compressed(
            os.path.join(OUT, f"msda_{name}.npz"),
            # inputs are fp32-representable: store them as float32 (exact), outputs as float64
            value=value.float().numpy(), spatial_shapes=shapes.numpy(), level_start_index=lsi.numpy(),
            sampling_loc=loc.float().numpy(), attn_weight=attn.float().numpy(),
            grad_out=grad_out.float().numpy(),
            out=out.detach().numpy(), grad_value=v.grad.numpy(), grad_sampling_loc=l_.grad.numpy(),
            grad_attn_weight=a.grad.numpy())
        print(f"{name}: out {tuple(out.shape)} |out|max {out.abs().max():.3e}")


if __name__ == "__main__":
    main()
