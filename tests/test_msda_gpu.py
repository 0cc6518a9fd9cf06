"""GPU parity tests of the MSDeformAttn kernels, called through the C ABI
(rlipv2_b200/dropin/MultiScaleDeformableAttention.py -> include/rlipv2_msda.h).

Checker = the CPU oracle (oracle/msda_oracle.c) and the golden fixtures produced by the reference's
own `ms_deform_attn_core_pytorch`.  Tolerances: north_star says 1e-3 rel fp32; the reference's own
test uses rtol 1e-2 / atol 1e-3 for fp32 and torch.allclose defaults for fp64
(models/ops/test.py:44,60).  We hold fp32 to rtol 1e-4 (+ a small atol for near-zero sums)."""
import numpy as np
import pytest
import torch

from oracle import msda_oracle
from tests.golden_util import load_msda, msda_cases

pytestmark = pytest.mark.gpu


def _msda():
    from rlipv2_b200.dropin import MultiScaleDeformableAttention as MSDA
    return MSDA


def _dev(g, dtype):
    f = lambda k: torch.from_numpy(g[k]).to("cuda", dtype).contiguous()
    return (f("value"), torch.from_numpy(g["spatial_shapes"]).cuda(),
            torch.from_numpy(g["level_start_index"]).cuda(), f("sampling_loc"), f("attn_weight"),
            f("grad_out"))


@pytest.mark.parametrize("name", msda_cases())
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_golden(name, dtype):
    g = load_msda(name)
    value, shapes, lsi, loc, attn, gout = _dev(g, dtype)
    out = _msda().ms_deform_attn_forward(value, shapes, lsi, loc, attn, 64)
    gv, gl, ga = _msda().ms_deform_attn_backward(value, shapes, lsi, loc, attn, gout, 64)
    if dtype == torch.float64:
        tol = dict(rtol=1e-7, atol=1e-11)
        ltol = dict(rtol=1e-6, atol=1e-10)
    else:
        tol = dict(rtol=1e-4, atol=1e-6)
        ltol = dict(rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(out.cpu().numpy(), g["out"], **tol)
    np.testing.assert_allclose(gv.cpu().numpy(), g["grad_value"], **tol)
    np.testing.assert_allclose(gl.cpu().numpy(), g["grad_sampling_loc"], **ltol)
    np.testing.assert_allclose(ga.cpu().numpy(), g["grad_attn_weight"], **tol)


def _rand_inputs(N, Lq, M, D, shapes, seed, lo=0.0, hi=1.0, L=None, P=4):
    # models/ops/test.py:37-40 recipe
    g = torch.Generator().manual_seed(seed)
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    value = torch.rand(N, S, M, D, generator=g) * 0.01
    loc = torch.rand(N, Lq, M, L, P, 2, generator=g) * (hi - lo) + lo
    attn = torch.rand(N, Lq, M, L, P, generator=g) + 1e-5
    attn = attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    gout = torch.randn(N, Lq, M * D, generator=g)
    sh = torch.as_tensor(shapes, dtype=torch.long)
    lsi = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
    return value, sh, lsi, loc, attn, gout


@pytest.mark.parametrize("cfg", [
    # (N, Lq, M, D, shapes, lo, hi)
    (2, 300, 8, 32, [(100, 100), (50, 50), (25, 25), (13, 13)], 0.0, 1.0),      # config 5, decoder shaped
    (3, 77, 8, 32, [(25, 42), (13, 21), (7, 11), (4, 6)], -0.2, 1.2),           # ragged + out of range
    (1, 1657, 8, 32, [(31, 40), (16, 20), (8, 10), (4, 5)], 0.0, 1.0),          # encoder shaped, Lq == S
    (2, 33, 5, 32, [(9, 9), (5, 5), (3, 3), (2, 2)], 0.0, 1.0),                 # odd head count, tail CTA
    (2, 50, 4, 64, [(20, 20), (10, 10), (5, 5), (3, 3)], 0.0, 1.0),             # generic path D=64
    (1, 40, 2, 30, [(20, 20), (10, 10)], 0.0, 1.0),                             # generic path, L=2
])
def test_against_oracle_fp32(cfg):
    N, Lq, M, D, shapes, lo, hi = cfg
    value, sh, lsi, loc, attn, gout = _rand_inputs(N, Lq, M, D, shapes, seed=3, lo=lo, hi=hi)
    ref_out = msda_oracle.forward(value.numpy(), sh.numpy(), lsi.numpy(), loc.numpy(), attn.numpy())
    rgv, rgl, rga = msda_oracle.backward(value.numpy(), sh.numpy(), lsi.numpy(), loc.numpy(),
                                         attn.numpy(), gout.numpy())
    c = lambda t: t.cuda().contiguous()
    out = _msda().ms_deform_attn_forward(c(value), c(sh), c(lsi), c(loc), c(attn), 64)
    gv, gl, ga = _msda().ms_deform_attn_backward(c(value), c(sh), c(lsi), c(loc), c(attn), c(gout), 64)
    np.testing.assert_allclose(out.cpu().numpy(), ref_out, rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(ga.cpu().numpy(), rga, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(gl.cpu().numpy(), rgl, rtol=1e-3, atol=2e-5)
    # grad_value: many fp32 contributions per cell, summed in a different (atomic) order
    np.testing.assert_allclose(gv.cpu().numpy(), rgv, rtol=1e-3, atol=1e-5)


def test_empty_query_and_empty_batch():
    value, sh, lsi, loc, attn, gout = _rand_inputs(2, 0, 8, 32, [(4, 4), (2, 2), (1, 1), (1, 1)], 1)
    c = lambda t: t.cuda().contiguous()
    out = _msda().ms_deform_attn_forward(c(value), c(sh), c(lsi), c(loc), c(attn), 64)
    assert out.shape == (2, 0, 256)
    gv, gl, ga = _msda().ms_deform_attn_backward(c(value), c(sh), c(lsi), c(loc), c(attn), c(gout), 64)
    assert gv.shape == value.shape and float(gv.abs().sum()) == 0.0
    assert gl.shape == loc.shape and ga.shape == attn.shape


def test_precondition_errors_match_reference():
    value, sh, lsi, loc, attn, gout = _rand_inputs(3, 5, 8, 32, [(4, 4), (2, 2), (1, 1), (1, 1)], 1)
    c = lambda t: t.cuda().contiguous()
    with pytest.raises(RuntimeError, match="contiguous"):        # ms_deform_attn_cuda.cu:28
        _msda().ms_deform_attn_forward(c(value).transpose(2, 3).contiguous().transpose(2, 3),
                                       c(sh), c(lsi), c(loc), c(attn), 64)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):   # cu:35
        _msda().ms_deform_attn_forward(c(value), sh, c(lsi), c(loc), c(attn), 64)
    with pytest.raises(RuntimeError, match="must divide"):        # cu:52
        _msda().ms_deform_attn_forward(c(value), c(sh), c(lsi), c(loc), c(attn), 2)


def test_autograd_function_and_gradcheck_fp64():
    # models/ops/test.py:67-90: gradcheck in fp64 over the channel list (one per dispatch branch)
    from rlipv2_b200.ms_deform_attn import MSDeformAttnFunction
    from torch.autograd import gradcheck
    N, M, Lq, L, P = 1, 2, 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long).cuda()
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    torch.manual_seed(3)
    for channels in [30, 32, 64, 71]:
        value = (torch.rand(N, S, M, channels).cuda() * 0.01).double().requires_grad_(True)
        loc = torch.rand(N, Lq, M, L, P, 2).cuda().double().requires_grad_(True)
        attn = torch.rand(N, Lq, M, L, P).cuda() + 1e-5
        attn = (attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)).double().requires_grad_(True)
        assert gradcheck(MSDeformAttnFunction.apply, (value, shapes, lsi, loc, attn, 2))


def test_full_size_properties():
    """BASELINE config sizes (encoder call of a 800x1333 image, S = Lq = 22223, batch 2) through
    size-independent properties: linearity in `value`, the adjoint identity
    <grad_out, f(v)> = <grad_value, v>, and constant-field reproduction."""
    shapes = [(100, 167), (50, 84), (25, 42), (13, 21)]
    N, M, D = 2, 8, 32
    S = sum(h * w for h, w in shapes)
    g = torch.Generator(device="cuda").manual_seed(5)
    dev = "cuda"
    sh = torch.as_tensor(shapes, dtype=torch.long, device=dev)
    lsi = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
    v1 = torch.randn(N, S, M, D, device=dev, generator=g)
    v2 = torch.randn(N, S, M, D, device=dev, generator=g)
    # interior samples only (full bilinear support): loc in [0.1, 0.9]
    loc = torch.rand(N, S, M, 4, 4, 2, device=dev, generator=g) * 0.8 + 0.1
    attn = torch.softmax(torch.randn(N, S, M, 16, device=dev, generator=g), -1).view(N, S, M, 4, 4)
    gout = torch.randn(N, S, M * D, device=dev, generator=g)
    F = lambda v: _msda().ms_deform_attn_forward(v, sh, lsi, loc, attn, 64)
    o1, o2 = F(v1), F(v2)
    o12 = F(2.0 * v1 - 3.0 * v2)
    torch.testing.assert_close(o12, 2.0 * o1 - 3.0 * o2, rtol=1e-4, atol=1e-4)
    oc = F(torch.full_like(v1, 1.5))
    torch.testing.assert_close(oc, torch.full_like(oc, 1.5), rtol=1e-5, atol=1e-5)
    gv, gl, ga = _msda().ms_deform_attn_backward(v1, sh, lsi, loc, attn, gout, 64)
    lhs = (gout.double() * o1.double()).sum()
    rhs = (gv.double() * v1.double()).sum()
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), 1.0) + 1e-2 * float(gout.numel()) ** 0.5 * 1e-3
    # d out / d attn = sampled value: <ga, attn> over points equals <gout, out> per pair
    lhs2 = (ga.double() * attn.double()).sum()
    assert abs(lhs2 - lhs) <= 1e-4 * max(abs(lhs), 1.0) + 1e-1
    assert torch.isfinite(gl).all()


def test_against_reference_cuda_kernel():
    """Second oracle: the reference's OWN CUDA op, compiled unmodified from /root/reference into
    oracle/_ref/ (oracle/build_ref.py).  Skipped when that build did not travel."""
    from oracle import build_ref
    ref = build_ref.load()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    from rlipv2_b200 import synth
    for make in (lambda: synth.random_inputs(2, 300, synth.LEVELS_MICRO, seed=3),
                 lambda: synth.encoder_inputs(1, [(31, 40), (16, 20), (8, 10), (4, 5)], seed=4)):
        value, sh, lsi, loc, attn, gout = make()
        out = _msda().ms_deform_attn_forward(value, sh, lsi, loc, attn, 64)
        rout = ref.ms_deform_attn_forward(value, sh, lsi, loc, attn, 64)
        torch.testing.assert_close(out, rout, rtol=1e-4, atol=1e-5)
        gv, gl, ga = _msda().ms_deform_attn_backward(value, sh, lsi, loc, attn, gout, 64)
        rgv, rgl, rga = ref.ms_deform_attn_backward(value, sh, lsi, loc, attn, gout, 64)
        torch.testing.assert_close(ga, rga, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(gl, rgl, rtol=1e-3, atol=1e-2)
        torch.testing.assert_close(gv, rgv, rtol=1e-3, atol=1e-3)


def test_full_size_against_reference_cuda_kernel():
    """the reference's own CUDA op at the BASELINE size - the encoder call of a 800x1333 image,
    N = 2, S = Lq = 22223 (models/ops/src/cuda/ms_deform_im2col_cuda.cuh:238-403) - not only at Lq = 300."""
    from oracle import build_ref
    ref = build_ref.load()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    from rlipv2_b200 import synth
    value, sh, lsi, loc, attn, gout = synth.encoder_inputs(2, synth.LEVELS_800x1333, seed=7)
    out = _msda().ms_deform_attn_forward(value, sh, lsi, loc, attn, 64)
    rout = ref.ms_deform_attn_forward(value, sh, lsi, loc, attn, 64)
    torch.testing.assert_close(out, rout, rtol=1e-4, atol=2e-5)
    gv, gl, ga = _msda().ms_deform_attn_backward(value, sh, lsi, loc, attn, gout, 64)
    rgv, rgl, rga = ref.ms_deform_attn_backward(value, sh, lsi, loc, attn, gout, 64)
    torch.testing.assert_close(ga, rga, rtol=1e-4, atol=1e-4)
    # grad_value / grad_loc are sums of up to hundreds of fp32 atomics in a different order on each side
    assert float((gv - rgv).abs().max()) <= 1e-3 * float(rgv.abs().max())
    assert float((gl - rgl).abs().max()) <= 1e-3 * float(rgl.abs().max())


def test_misaligned_views_take_the_generic_kernel():
    """a contiguous tensor whose storage offset is not 16-byte aligned (the reference's scalar kernel accepts
    it: only is_contiguous() is checked, ms_deform_attn_cuda.cu:28-33) must not reach the 128-bit fast path."""
    from rlipv2_b200 import synth
    shapes = [(12, 15), (6, 8), (3, 4), (2, 2)]
    value, sh, lsi, loc, attn, gout = synth.random_inputs(2, 37, shapes, seed=5)

    def shifted(t):
        buf = torch.empty(t.numel() + 1, device=t.device, dtype=t.dtype)
        v = buf[1:].view(t.shape)
        v.copy_(t)
        assert v.is_contiguous() and v.data_ptr() % 16 != 0
        return v

    out = _msda().ms_deform_attn_forward(value, sh, lsi, loc, attn, 64)
    gv, gl, ga = _msda().ms_deform_attn_backward(value, sh, lsi, loc, attn, gout, 64)
    for which in range(4):
        args = [value, loc, attn, gout]
        args[which] = shifted(args[which])
        v, l, a, g = args
        o2 = _msda().ms_deform_attn_forward(v, sh, lsi, l, a, 64)
        torch.testing.assert_close(o2, out, rtol=1e-5, atol=1e-7)
        gv2, gl2, ga2 = _msda().ms_deform_attn_backward(v, sh, lsi, l, a, g, 64)
        torch.testing.assert_close(gv2, gv, rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(gl2, gl, rtol=1e-3, atol=1e-5)
        torch.testing.assert_close(ga2, ga, rtol=1e-4, atol=1e-6)


def test_tma_staged_variant_equals_the_default_kernel():
    """north_star's literal design (coarse levels staged in shared memory by TMA, rlipv2_msda_forward_tma_f32) against the
    default forward: same gather, same arithmetic per corner -> equal up to the order of the 64 fused multiply-adds."""
    from rlipv2_b200 import msda_abi, synth
    for shapes, N in ((synth.LEVELS_800x1333, 2), ([(31, 40), (16, 20), (8, 10), (4, 5)], 3)):
        value, sh, lsi, loc, attn, _ = synth.encoder_inputs(N, shapes, seed=11)
        coarse = sum(h * w for h, w in shapes[:2])
        ref = _msda().ms_deform_attn_forward(value, sh, lsi, loc, attn, 64)
        out = msda_abi.forward_tma(value, sh, lsi, loc, attn, coarse)
        torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)
    value, sh, lsi, loc, attn, _ = synth.random_inputs(2, 301, synth.LEVELS_MICRO, seed=3)      # out-of-range samples, ragged tail
    loc = loc * 1.4 - 0.2
    ref = _msda().ms_deform_attn_forward(value, sh, lsi, loc, attn, 64)
    out = msda_abi.forward_tma(value, sh, lsi, loc, attn, 100 * 100 + 50 * 50)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-6)
