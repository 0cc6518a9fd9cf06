"""GPU parity of the "paired" MSDeformAttn backward schedule (rlipv2_msda_set_backward_variant, include/rlipv2_msda.h):
two consecutive queries per lane group, corner loads shared and grad_value reductions merged where the bilinear
footprints coincide.  Checkers: the CPU oracle (oracle/msda_oracle.c, restating
/root/reference/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:302-403) on sampling patterns that exercise every merge
case, and the one-query-per-group schedule (itself oracle-checked in tests/test_msda_gpu.py) at full size."""
import numpy as np
import pytest
import torch

from oracle import msda_oracle

pytestmark = pytest.mark.gpu

SMALL = [(31, 40), (16, 20), (8, 10), (4, 5)]


@pytest.fixture
def paired_everywhere():
    from rlipv2_b200 import msda_abi
    keep = msda_abi.backward_variant()
    msda_abi.set_backward_variant(2)
    yield msda_abi
    msda_abi.set_backward_variant(keep)


def _encoder_like(N, shapes, M, noise_px, seed, quantise=False):
    """one query per cell in raster order; offsets = ring pattern + noise (in cells).  noise 0: neighbouring queries
    share / shift footprints (the merge cases); quantise: offsets on a 1/2-cell lattice, so many fractional parts are
    exactly 0 and many corner pairs tie"""
    from rlipv2_b200 import synth
    value, sh, lsi, loc, attn, gout = synth.encoder_inputs(N, shapes, M=M, seed=seed, noise_px=noise_px, device="cpu")
    if quantise:
        norm = torch.as_tensor([[w, h] for h, w in shapes], dtype=torch.float32).view(1, 1, 1, len(shapes), 1, 2)
        loc = torch.round(loc * norm * 2) / 2 / norm
    return value, sh, lsi, loc.contiguous(), attn, gout


@pytest.mark.parametrize("case", ["init_like", "noisy", "lattice", "random", "out_of_range", "one_cell_levels"])
def test_paired_backward_matches_oracle(paired_everywhere, case):
    from rlipv2_b200.dropin import MultiScaleDeformableAttention as MSDA
    from tests.test_msda_gpu import _rand_inputs
    if case == "init_like":
        value, sh, lsi, loc, attn, gout = _encoder_like(2, SMALL, 8, 0.0, 1)          # Lq = 1657 (odd: a pair spans images)
    elif case == "noisy":
        value, sh, lsi, loc, attn, gout = _encoder_like(1, SMALL, 8, 1.0, 2)
    elif case == "lattice":
        value, sh, lsi, loc, attn, gout = _encoder_like(2, SMALL, 5, 0.0, 3, quantise=True)
    elif case == "random":
        value, sh, lsi, loc, attn, gout = _rand_inputs(2, 301, 8, 32, [(25, 42), (13, 21), (7, 11), (4, 6)], seed=3)
    elif case == "out_of_range":
        value, sh, lsi, loc, attn, gout = _rand_inputs(3, 77, 8, 32, [(25, 42), (13, 21), (7, 11), (4, 6)], seed=4,
                                                       lo=-0.3, hi=1.3)
    else:                                                                             # 1-wide / 1-high levels: clamped corners
        value, sh, lsi, loc, attn, gout = _rand_inputs(2, 64, 4, 32, [(5, 1), (1, 7), (1, 1), (2, 2)], seed=5,
                                                       lo=-0.2, hi=1.2)
    rgv, rgl, rga = msda_oracle.backward(value.numpy(), sh.numpy(), lsi.numpy(), loc.numpy(), attn.numpy(), gout.numpy())
    c = lambda t: t.cuda().contiguous()
    gv, gl, ga = MSDA.ms_deform_attn_backward(c(value), c(sh), c(lsi), c(loc), c(attn), c(gout), 64)
    scale = float(np.abs(rgv).max())
    np.testing.assert_allclose(ga.cpu().numpy(), rga, rtol=1e-4, atol=1e-5 * float(np.abs(rga).max()))
    np.testing.assert_allclose(gl.cpu().numpy(), rgl, rtol=1e-3, atol=1e-5 * float(np.abs(rgl).max()))
    np.testing.assert_allclose(gv.cpu().numpy(), rgv, rtol=1e-3, atol=1e-5 * scale)


@pytest.mark.parametrize("ref_dim", [2, 4])
def test_paired_fused_prologue_backward_equals_unpaired(ref_dim):
    from rlipv2_b200 import msda_abi
    from tests.test_msda_proj_gpu import SHAPES, _inputs
    value, sh, lsi, ref, proj, gout = _inputs(2, 727, 8, SHAPES, 7, ref_dim=ref_dim, spread=0.5)
    dv = lambda t: t.cuda().contiguous()
    value, sh, lsi, ref, proj, gout = map(dv, (value, sh, lsi, ref, proj, gout))
    keep = msda_abi.backward_variant()
    res = []
    try:
        for variant in (0, 2):
            msda_abi.set_backward_variant(variant)
            gv = torch.empty_like(value)
            gp = torch.full_like(proj, float("nan"))
            msda_abi.proj_backward(value, sh, lsi, ref, proj, gout, gv, gp)
            res.append((gv, gp))
    finally:
        msda_abi.set_backward_variant(keep)
    (gv0, gp0), (gv1, gp1) = res
    assert torch.isfinite(gp1).all()
    torch.testing.assert_close(gp1, gp0, rtol=1e-4, atol=1e-5 * float(gp0.abs().max()))
    torch.testing.assert_close(gv1, gv0, rtol=1e-3, atol=1e-5 * float(gv0.abs().max()))


def test_paired_full_size_encoder_call_equals_unpaired():
    """BASELINE config 2 encoder call (S = Lq = 22223, batch 2): variant 1 against variant 0, plus the adjoint identity
    <grad_out, f(v)> = <grad_value, v> as a size-independent property"""
    from rlipv2_b200 import msda_abi, synth
    from rlipv2_b200.dropin import MultiScaleDeformableAttention as MSDA
    keep = msda_abi.backward_variant()
    try:
        for noise in (0.0, 1.0):
            value, sh, lsi, loc, attn, gout = synth.encoder_inputs(2, synth.LEVELS_800x1333, seed=6, noise_px=noise)
            msda_abi.set_backward_variant(0)
            gv0, gl0, ga0 = MSDA.ms_deform_attn_backward(value, sh, lsi, loc, attn, gout, 64)
            msda_abi.set_backward_variant(1)
            gv1, gl1, ga1 = MSDA.ms_deform_attn_backward(value, sh, lsi, loc, attn, gout, 64)
            torch.testing.assert_close(ga1, ga0, rtol=1e-4, atol=1e-5 * float(ga0.abs().max()))
            torch.testing.assert_close(gl1, gl0, rtol=1e-3, atol=1e-5 * float(gl0.abs().max()))
            torch.testing.assert_close(gv1, gv0, rtol=1e-3, atol=2e-5 * float(gv0.abs().max()))
            out = MSDA.ms_deform_attn_forward(value, sh, lsi, loc, attn, 64)
            lhs = (gout.double() * out.double()).sum()
            rhs = (gv1.double() * value.double()).sum()
            assert abs(lhs - rhs) <= 1e-4 * max(abs(float(lhs)), 1.0) + 1.0
    finally:
        msda_abi.set_backward_variant(keep)
