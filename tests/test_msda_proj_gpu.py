"""GPU parity of the fused-prologue MSDeformAttn entry points (rlipv2_msda_proj_*, include/rlipv2_msda.h).

Checker 1: the CPU oracle (oracle/msda_oracle.c) fed with locations / attention weights derived on the CPU
           exactly as /root/reference/models/ops/modules/ms_deform_attn.py:102-109 derives them, and torch
           autograd through that derivation for the gradient w.r.t. the raw projection.
Checker 2: the plain CUDA op behind the reference's own module arithmetic (same device, same inputs).
Tolerance: 1e-3 relative (north_star), held tighter where the arithmetic is identical."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import msda_oracle

pytestmark = pytest.mark.gpu

SHAPES = [(20, 27), (10, 14), (5, 7), (3, 4)]


def _inputs(N, Lq, M, shapes, seed, spread=3.0, ref_dim=2):
    g = torch.Generator().manual_seed(seed)
    S = sum(h * w for h, w in shapes)
    value = torch.randn(N, S, M, 32, generator=g)
    ref = torch.rand(N, Lq, 4, 2, generator=g) * 1.2 - 0.1                 # some points fall outside the map
    if ref_dim == 4:                                                       # anchors (cx, cy, w, h) per level
        ref = torch.cat((ref, torch.rand(N, Lq, 4, 2, generator=g) * 0.6 + 0.02), -1)
        spread = 1.5
    proj = torch.cat((torch.randn(N, Lq, M * 32, generator=g) * spread,    # raw offsets, in cells
                      torch.randn(N, Lq, M * 16, generator=g) * 2.0), -1)  # raw logits
    gout = torch.randn(N, Lq, M * 32, generator=g)
    sh = torch.as_tensor(shapes, dtype=torch.long)
    lsi = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
    return value, sh, lsi, ref, proj, gout


def _module_arithmetic(ref, proj, sh, M):
    """ms_deform_attn.py:102-109 on any device"""
    N, Lq = proj.shape[:2]
    off = proj[..., :M * 32].reshape(N, Lq, M, 4, 4, 2)
    attn = F.softmax(proj[..., M * 32:].reshape(N, Lq, M, 16), -1).view(N, Lq, M, 4, 4)
    if ref.shape[-1] == 4:                                                 # ms_deform_attn.py:110-112
        loc = ref[:, :, None, :, None, :2] + off / 4 * ref[:, :, None, :, None, 2:] * 0.5
        return loc.contiguous(), attn.contiguous()
    normalizer = torch.stack([sh[..., 1], sh[..., 0]], -1).to(proj.dtype)
    loc = ref[:, :, None, :, None, :] + off / normalizer[None, None, None, :, None, :]
    return loc.contiguous(), attn.contiguous()


@pytest.mark.parametrize("ref_dim", [2, 4])
@pytest.mark.parametrize("N,Lq,M,shapes,seed", [
    (2, 50, 8, SHAPES, 1), (1, 1, 8, SHAPES, 2), (3, 37, 2, SHAPES, 3),
    (2, sum(h * w for h, w in SHAPES), 8, SHAPES, 4),                       # encoder-shaped: Lq == S
    (1, 9000, 8, [(64, 80), (32, 40), (16, 20), (8, 10)], 5),              # NQ >= 8192 kernel variant
])
def test_proj_matches_oracle_and_plain_op(N, Lq, M, shapes, seed, ref_dim):
    from rlipv2_b200 import msda_abi
    from rlipv2_b200.dropin import MultiScaleDeformableAttention as MSDA
    value, sh, lsi, ref, proj, gout = _inputs(N, Lq, M, shapes, seed, ref_dim=ref_dim)
    dv = lambda t: t.cuda().contiguous()
    value_d, sh_d, lsi_d, ref_d, proj_d, gout_d = map(dv, (value, sh, lsi, ref, proj, gout))
    out = torch.empty(N, Lq, M * 32, device="cuda")
    msda_abi.proj_forward(value_d, sh_d, lsi_d, ref_d, proj_d, out)
    gv = torch.empty_like(value_d)
    gp = torch.full_like(proj_d, float("nan"))                            # every element must be written
    msda_abi.proj_backward(value_d, sh_d, lsi_d, ref_d, proj_d, gout_d, gv, gp)
    assert torch.isfinite(gp).all()

    # checker 2: plain CUDA op + the module's arithmetic on the device, autograd for the chain rule
    proj_r = proj_d.clone().requires_grad_(True)
    loc_d, attn_d = _module_arithmetic(ref_d, proj_r, sh_d, M)
    out2 = MSDA.ms_deform_attn_forward(value_d, sh_d, lsi_d, loc_d.detach().contiguous(), attn_d.detach().contiguous(), 64)
    gv2, gl2, ga2 = MSDA.ms_deform_attn_backward(value_d, sh_d, lsi_d, loc_d.detach().contiguous(),
                                                 attn_d.detach().contiguous(), gout_d, 64)
    torch.autograd.backward([loc_d, attn_d], [gl2, ga2])
    scale = lambda t: float(t.abs().max()) + 1e-12
    assert float((out - out2).abs().max()) <= 2e-5 * scale(out2)
    assert float((gv - gv2).abs().max()) <= 1e-4 * scale(gv2)
    assert float((gp - proj_r.grad).abs().max()) <= 2e-4 * scale(proj_r.grad)

    # checker 1: CPU oracle on CPU-derived locations / weights (skip the two large cases: seconds matter)
    if N * Lq * M <= 20000:
        proj_c = proj.clone().requires_grad_(True)
        loc_c, attn_c = _module_arithmetic(ref, proj_c, sh, M)
        c = lambda t: t.detach().numpy()
        ref_out = msda_oracle.forward(c(value), c(sh), c(lsi), c(loc_c), c(attn_c))
        rgv, rgl, rga = msda_oracle.backward(c(value), c(sh), c(lsi), c(loc_c), c(attn_c), c(gout))
        torch.autograd.backward([loc_c, attn_c], [torch.from_numpy(rgl), torch.from_numpy(rga)])
        np.testing.assert_allclose(out.cpu().numpy(), ref_out, rtol=1e-3, atol=1e-4 * scale(out2))
        np.testing.assert_allclose(gv.cpu().numpy(), rgv, rtol=1e-3, atol=1e-4 * scale(gv2))
        np.testing.assert_allclose(gp.cpu().numpy(), c(proj_c.grad), rtol=1e-3, atol=2e-4 * scale(proj_r.grad))


@pytest.mark.parametrize("ref_dim", [2, 4])
def test_module_fused_prologue_equals_unfused(monkeypatch, ref_dim):
    """MSDeformAttn.forward with 2-d reference points (encoder) and detached 4-d anchors (decoder layers >= 1):
    fused-prologue path == reference arithmetic path"""
    import rlipv2_b200.ms_deform_attn as mod
    from rlipv2_b200 import dense
    dense.set_matmul_precision("fp32")
    torch.manual_seed(0)
    S = sum(h * w for h, w in SHAPES)
    m = mod.MSDeformAttn(256, 4, 8, 4).cuda()
    with torch.no_grad():
        m.sampling_offsets.weight.normal_(0, 0.02)
        m.attention_weights.weight.normal_(0, 0.05)
    src = torch.randn(2, S, 256, device="cuda")
    refp = torch.rand(2, S, 4, 2, device="cuda")
    if ref_dim == 4:
        refp = torch.cat((refp, torch.rand(2, S, 4, 2, device="cuda") * 0.5 + 0.02), -1)
    sh = torch.as_tensor(SHAPES, dtype=torch.long, device="cuda")
    lsi = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
    mask = torch.zeros(2, S, dtype=torch.bool, device="cuda")
    mask[1, -40:] = True
    res = {}
    for fused in (False, True):
        monkeypatch.setattr(mod, "_FUSED_PROLOGUE", fused)
        x = src.clone().requires_grad_(True)
        for p in m.parameters():
            p.grad = None
        y = m(x + 0.1, refp, x, sh, lsi, mask)
        (y * torch.linspace(-1, 1, 256, device="cuda")).sum().backward()
        res[fused] = (y.detach(), x.grad.detach(), {n: p.grad.detach().clone() for n, p in m.named_parameters()})
    a, b = res[False], res[True]
    rel = lambda u, v: float((u - v).abs().max() / (v.abs().max() + 1e-12))
    assert rel(b[0], a[0]) < 1e-4 and rel(b[1], a[1]) < 1e-3
    for n in a[2]:
        assert rel(b[2][n], a[2][n]) < 1e-3, n
