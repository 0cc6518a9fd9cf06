"""GPU parity of the merged-reduction backward of the fast MSDeformAttn path (rlipv2_msda_set_backward_mode(1),
include/rlipv2_msda.h; rule in rlipv2_b200/csrc/msda_merge.h, checked on the host by tests/test_msda_merge_core.py).

Checkers: the reference's golden fixtures, the CPU oracle (oracle/msda_oracle.c), the reference's own CUDA op at the
BASELINE size (oracle/_ref) and the unmerged schedule of this library (mode 0) on the same inputs.  grad_value is a sum of
fp32 contributions in another order in every one of them: 1e-3 of its scale (north_star), much tighter against mode 0;
grad_sampling_loc / grad_attn_weight do not depend on the mode (same expressions; 1e-6 of scale allows for another
fused-multiply-add contraction in the other instantiation)."""
import numpy as np
import pytest
import torch

from oracle import msda_oracle
from tests.golden_util import load_msda, msda_cases

pytestmark = pytest.mark.gpu


def _msda():
    from rlipv2_b200.dropin import MultiScaleDeformableAttention as MSDA
    return MSDA


@pytest.fixture
def modes():
    from rlipv2_b200 import msda_abi
    before = msda_abi.get_backward_mode()
    yield msda_abi.set_backward_mode
    msda_abi.set_backward_mode(before)


def _same(a, b):
    return float((a - b).abs().max()) <= 1e-6 * float(b.abs().max()) + 1e-30


def _both(modes, fn):
    modes(0)
    a = fn()
    modes(1)
    b = fn()
    return a, b


def test_mode_switch_round_trip(modes):
    from rlipv2_b200 import msda_abi
    modes(1)
    assert msda_abi.get_backward_mode() == 1
    modes(0)
    assert msda_abi.get_backward_mode() == 0
    for m in (2, 3, 4):
        modes(m)
        assert msda_abi.get_backward_mode() == m
    with pytest.raises(RuntimeError):
        msda_abi.set_backward_mode(7)


@pytest.mark.parametrize("name", msda_cases())
def test_golden_fp32_merged(name, modes):
    g = load_msda(name)
    f = lambda k: torch.from_numpy(g[k]).to("cuda", torch.float32).contiguous()
    value, loc, attn, gout = f("value"), f("sampling_loc"), f("attn_weight"), f("grad_out")
    shapes, lsi = torch.from_numpy(g["spatial_shapes"]).cuda(), torch.from_numpy(g["level_start_index"]).cuda()
    modes(1)
    gv, gl, ga = _msda().ms_deform_attn_backward(value, shapes, lsi, loc, attn, gout, 64)
    np.testing.assert_allclose(gv.cpu().numpy(), g["grad_value"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(gl.cpu().numpy(), g["grad_sampling_loc"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(ga.cpu().numpy(), g["grad_attn_weight"], rtol=1e-4, atol=1e-6)


def _clustered(N, Lq, M, shapes, seed, noise, lo=0.0, hi=1.0):
    """reference points anywhere in [lo, hi]^2 (out-of-range included), the four points of a level whole cells apart along
    the head's direction + noise: corners coincide, partially overlap, sit on borders and on exact cell centres"""
    g = torch.Generator().manual_seed(seed)
    S = sum(h * w for h, w in shapes)
    value = torch.randn(N, S, M, 32, generator=g)
    ref = torch.rand(N, Lq, 1, 1, 1, 2, generator=g) * (hi - lo) + lo
    if noise == 0.0:                                               # snap to cell centres of level 0: fractional parts 0
        H0, W0 = shapes[0]
        ref = (torch.floor(ref * torch.tensor([W0, H0])) + 0.5) / torch.tensor([W0, H0])
    th = torch.arange(M) * (2.0 * np.pi / M)
    ring = torch.stack([th.cos(), th.sin()], -1)
    ring = ring / ring.abs().max(-1, keepdim=True)[0]
    offs = ring.view(1, 1, M, 1, 1, 2) * torch.arange(1, 5).view(1, 1, 1, 1, 4, 1)
    offs = offs + noise * torch.randn(N, Lq, M, 4, 4, 2, generator=g)
    norm = torch.as_tensor([[w, h] for h, w in shapes], dtype=torch.float32).view(1, 1, 1, 4, 1, 2)
    loc = (ref + offs / norm).contiguous()
    attn = torch.softmax(torch.randn(N, Lq, M, 16, generator=g), -1).view(N, Lq, M, 4, 4).contiguous()
    attn[:, ::7, :, 1, 2] = 0.0                                    # some exactly weightless points
    gout = torch.randn(N, Lq, M * 32, generator=g)
    sh = torch.as_tensor(shapes, dtype=torch.long)
    lsi = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
    return value, sh, lsi, loc, attn, gout


@pytest.mark.parametrize("cfg", [
    # (N, Lq, M, shapes, noise, lo, hi)
    (2, 300, 8, [(100, 100), (50, 50), (25, 25), (13, 13)], 0.5, 0.0, 1.0),
    (3, 77, 8, [(25, 42), (13, 21), (7, 11), (4, 6)], 0.3, -0.2, 1.2),          # ragged + out of range
    (1, 1657, 8, [(31, 40), (16, 20), (8, 10), (4, 5)], 0.0, 0.0, 1.0),         # whole-cell ring on cell centres
    (2, 33, 5, [(9, 9), (5, 5), (3, 3), (2, 2)], 1.0, 0.0, 1.0),                # odd head count, tail CTA
    (2, 61, 8, [(4, 4), (2, 2), (1, 1), (1, 1)], 0.2, -0.5, 1.5),               # 1x1 levels: everything collides / clamps
])
def test_merged_against_oracle_and_unmerged(cfg, modes):
    N, Lq, M, shapes, noise, lo, hi = cfg
    value, sh, lsi, loc, attn, gout = _clustered(N, Lq, M, shapes, seed=11, noise=noise, lo=lo, hi=hi)
    rgv, rgl, rga = msda_oracle.backward(value.numpy(), sh.numpy(), lsi.numpy(), loc.numpy(), attn.numpy(), gout.numpy())
    c = lambda t: t.cuda().contiguous()
    args = tuple(map(c, (value, sh, lsi, loc, attn, gout)))
    (gv0, gl0, ga0), (gv1, gl1, ga1) = _both(modes, lambda: _msda().ms_deform_attn_backward(*args, 64))
    scale = float(np.abs(rgv).max())
    assert float(np.abs(gv1.cpu().numpy() - rgv).max()) <= 1e-4 * scale
    assert float((gv1 - gv0).abs().max()) <= 2e-5 * scale
    assert _same(gl1, gl0) and _same(ga1, ga0)
    # (grad_attn / grad_loc do not depend on the mode - checked above - and are sums with cancellation: tolerance of their scale)
    np.testing.assert_allclose(ga1.cpu().numpy(), rga, rtol=1e-4, atol=1e-5 * float(np.abs(rga).max()))
    if noise > 0.0:
        # (on exact cell centres floor() of the sample position is decided by the last bit of `loc * W - 0.5` - fused on the
        # device, two roundings in the host oracle - and grad_loc, unlike the output and grad_value, is discontinuous there)
        np.testing.assert_allclose(gl1.cpu().numpy(), rgl, rtol=1e-3, atol=1e-3 * float(np.abs(rgl).max()))


@pytest.mark.parametrize("mode", [2, 3, 4])
def test_call_shape_mode_and_three_cta_variants_equal_mode0(mode, modes):
    """mode 2 runs decoder-shaped calls of the plain op on the 3-CTA-per-SM instantiation of the unmerged kernel and
    encoder-shaped calls on the merged one; 3 / 4 force the 3-CTA variants: same sums as mode 0 in every case"""
    from rlipv2_b200 import synth
    for args in (synth.random_inputs(2, 300, synth.LEVELS_MICRO, seed=3),
                 _dev_tuple(_clustered(3, 77, 8, [(25, 42), (13, 21), (7, 11), (4, 6)], seed=5, noise=0.3, lo=-0.2, hi=1.2)),
                 synth.encoder_inputs(1, [(70, 100), (35, 50), (18, 25), (9, 13)], seed=2, noise_px=0.5)):     # NQ = 9392
        run = lambda: _msda().ms_deform_attn_backward(*args[:5], args[5], 64)
        modes(0)
        gv0, gl0, ga0 = run()
        modes(mode)
        gv1, gl1, ga1 = run()
        assert float((gv1 - gv0).abs().max()) <= 2e-5 * float(gv0.abs().max())
        assert _same(gl1, gl0) and _same(ga1, ga0)


def _dev_tuple(t):
    return tuple(x.cuda().contiguous() for x in t)


def test_merged_random_locations_equal_unmerged(modes):
    """models/ops/test.py recipe (uniform random locations: almost nothing merges) - the merged schedule must not change it"""
    from rlipv2_b200 import synth
    args = synth.random_inputs(2, 300, synth.LEVELS_MICRO, seed=3)
    (gv0, gl0, ga0), (gv1, gl1, ga1) = _both(modes, lambda: _msda().ms_deform_attn_backward(*args[:5], args[5], 64))
    assert float((gv1 - gv0).abs().max()) <= 2e-5 * float(gv0.abs().max())
    assert _same(gl1, gl0) and _same(ga1, ga0)


@pytest.mark.parametrize("noise", [0.0, 1.0])
def test_merged_full_size_against_reference_cuda_kernel(noise, modes):
    """the encoder call of a 800x1333 image, N = 2, S = Lq = 22223, against the reference's own CUDA op
    (ms_deform_im2col_cuda.cuh:302-403) and against mode 0; noise 0 = what a random-init step samples"""
    from oracle import build_ref
    from rlipv2_b200 import synth
    value, sh, lsi, loc, attn, gout = synth.encoder_inputs(2, synth.LEVELS_800x1333, seed=7, noise_px=noise)
    (gv0, gl0, ga0), (gv1, gl1, ga1) = _both(modes, lambda: _msda().ms_deform_attn_backward(value, sh, lsi, loc, attn, gout, 64))
    assert float((gv1 - gv0).abs().max()) <= 2e-5 * float(gv0.abs().max())
    assert _same(gl1, gl0) and _same(ga1, ga0)
    # adjoint identity <grad_out, f(value)> = <grad_value, value>
    out = _msda().ms_deform_attn_forward(value, sh, lsi, loc, attn, 64)
    lhs, rhs = (gout.double() * out.double()).sum(), (gv1.double() * value.double()).sum()
    assert abs(lhs - rhs) <= 1e-4 * max(abs(float(lhs)), 1.0) + 1.0
    ref = build_ref.load()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    rgv, rgl, rga = ref.ms_deform_attn_backward(value, sh, lsi, loc, attn, gout, 64)
    assert float((gv1 - rgv).abs().max()) <= 1e-3 * float(rgv.abs().max())
    assert float((gl1 - rgl).abs().max()) <= 1e-3 * float(rgl.abs().max())
    torch.testing.assert_close(ga1, rga, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("ref_dim", [2, 4])
def test_merged_fused_prologue_backward_equals_unmerged(ref_dim, modes):
    """rlipv2_msda_proj_*_backward_f32 (encoder, 2-d reference points) and the 4-d anchor variant (decoders)"""
    from rlipv2_b200 import msda_abi
    from tests.test_msda_proj_gpu import SHAPES, _inputs
    for N, Lq, M, spread in ((2, sum(h * w for h, w in SHAPES), 8, 1.0), (3, 37, 2, 3.0), (2, 300, 8, 0.0)):
        value, sh, lsi, ref, proj, gout = _inputs(N, Lq, M, SHAPES, seed=4, spread=spread, ref_dim=ref_dim)
        if spread == 0.0:                                           # the initialisation: whole-cell ring offsets
            th = torch.arange(M) * (2.0 * np.pi / M)
            ring = torch.stack([th.cos(), th.sin()], -1)
            ring = ring / ring.abs().max(-1, keepdim=True)[0]
            off = (ring.view(M, 1, 1, 2) * torch.arange(1, 5).view(1, 1, 4, 1)).expand(M, 4, 4, 2)
            proj[..., :M * 32] = off.reshape(-1)
        dv = lambda t: t.cuda().contiguous()
        value_d, sh_d, lsi_d, ref_d, proj_d, gout_d = map(dv, (value, sh, lsi, ref, proj, gout))

        def run():
            gv = torch.empty_like(value_d)
            gp = torch.full_like(proj_d, float("nan"))
            msda_abi.proj_backward(value_d, sh_d, lsi_d, ref_d, proj_d, gout_d, gv, gp)
            return gv, gp
        (gv0, gp0), (gv1, gp1) = _both(modes, run)
        assert torch.isfinite(gp1).all()
        assert _same(gp1, gp0)
        assert float((gv1 - gv0).abs().max()) <= 2e-5 * float(gv0.abs().max())
