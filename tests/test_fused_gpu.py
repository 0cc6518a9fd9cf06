"""GPU tests of the fused HBM-bound kernels (csrc/fused_ops.cu) through their C ABI, against plain
PyTorch fp32/fp64 references of the same ops (these are floating-point kernels)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,C", [(1, 128), (37, 256), (4446, 256), (512, 768), (300, 1024)])
@pytest.mark.parametrize("with_r", [True, False])
def test_add_layernorm_fwd_bwd(M, C, with_r):
    from rlipv2_b200 import fused_abi
    g = torch.Generator(device="cuda").manual_seed(M * 7 + C)
    x = torch.randn(M, C, device="cuda", generator=g)
    r = torch.randn(M, C, device="cuda", generator=g) if with_r else None
    w = torch.randn(C, device="cuda", generator=g)
    b = torch.randn(C, device="cuda", generator=g)
    dy = torch.randn(M, C, device="cuda", generator=g)
    y, z, mean, rstd = fused_abi.add_layernorm_fwd(x, r, w, b, 1e-5)
    zr = (x + r if with_r else x).double().requires_grad_(True)
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = F.layer_norm(zr, (C,), wr, br, 1e-5)
    torch.testing.assert_close(y.double(), yr, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(z.double(), zr.detach(), rtol=0, atol=0)
    yr.backward(dy.double())
    dz, dgamma, dbeta = fused_abi.layernorm_bwd(dy, z, mean, rstd, w)
    torch.testing.assert_close(dz.double(), zr.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dgamma.double(), wr.grad, rtol=1e-4, atol=1e-3 * max(1.0, M ** 0.5))
    torch.testing.assert_close(dbeta.double(), br.grad, rtol=1e-4, atol=1e-3 * max(1.0, M ** 0.5))


@pytest.mark.parametrize("M,N", [(1, 32), (100, 256), (44446, 2048), (513, 128)])
@pytest.mark.parametrize("mask", [True, False])
def test_relu_bwd_colsum(M, N, mask):
    from rlipv2_b200 import fused_abi
    g = torch.randn(M, N, device="cuda")
    y = torch.relu(torch.randn(M, N, device="cuda")) if mask else None
    gm, cs = fused_abi.relu_bwd_colsum(g.clone(), y)
    ref = g * (y > 0) if mask else g
    torch.testing.assert_close(gm, ref, rtol=0, atol=0)
    torch.testing.assert_close(cs.double(), ref.double().sum(0), rtol=1e-4, atol=1e-3 * max(1.0, M ** 0.5))


def test_flat_adamw_matches_torch():
    from rlipv2_b200 import fused_abi
    n = 1_000_003
    torch.manual_seed(0)
    p0 = torch.randn(n + 1, device="cuda")[:n]                 # odd length exercises the scalar tail
    p0 = p0.clone()
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref_p], lr=1.41e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)
    p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    step = torch.zeros((), device="cuda")
    for it in range(4):
        grad = torch.randn(n, device="cuda") * (0.1 + it)
        ref_p.grad = grad.clone()
        opt.step()
        step += 1
        fused_abi.adamw(p, grad, m, v, 1.41e-4, 0.9, 0.999, 1e-8, 1e-4, step)
    torch.testing.assert_close(p, ref_p.detach(), rtol=1e-5, atol=1e-6)
    st = opt.state[ref_p]
    # moments are O(1) sums of O(1) terms that can cancel: a few ulps of 1.0 (1.2e-7) absolute, fma vs mul+add
    torch.testing.assert_close(m, st["exp_avg"], rtol=1e-5, atol=5e-7)
    torch.testing.assert_close(v, st["exp_avg_sq"], rtol=1e-5, atol=1e-9)


def test_flat_adamw_grad_scale_equals_prescaled_gradient():
    """rlipv2_adamw_scaled_f32: the clip coefficient / rank-mean factor applied inside the optimizer's gradient read"""
    from rlipv2_b200 import fused_abi
    n = 300_001
    torch.manual_seed(1)
    p0 = torch.randn(n, device="cuda")
    grad = torch.randn(n, device="cuda") * 3
    scale = torch.tensor([0.0371], device="cuda")
    step = torch.ones((), device="cuda")
    res = []
    for fused in (True, False):
        p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
        if fused:
            fused_abi.adamw(p, grad, m, v, 1e-3, 0.9, 0.999, 1e-8, 1e-4, step, grad_scale=scale)
        else:
            fused_abi.adamw(p, grad * scale, m, v, 1e-3, 0.9, 0.999, 1e-8, 1e-4, step)
        res.append((p, m, v))
    for a, b in zip(*res):
        assert torch.equal(a, b)          # the same fp32 product, formed in the kernel instead of by torch


@pytest.mark.parametrize("channels_last", [True, False])
@pytest.mark.parametrize("fuse", [False, True])
def test_group_norm_tokens_matches_torch(channels_last, fuse):
    """dense.group_norm_tokens (csrc/fused_ops.cu gn_tok_*): GroupNorm(32, 256) of every feature level written as rows
    of one [N, sum HW, 256] token tensor == F.group_norm + flatten(2).transpose(1, 2) + cat (the reference's input
    projections + dab_deformable/deformable_transformer.py:452-470), forward and backward"""
    from rlipv2_b200 import dense
    torch.manual_seed(3)
    N = 2
    sizes = [(25, 42), (13, 21), (7, 11), (1, 3)]             # HW = 1050, 273, 77, 3: chunk tails, tiny level
    norms = [torch.nn.GroupNorm(32, 256).cuda() for _ in sizes]
    for n in norms:
        with torch.no_grad():
            n.weight.uniform_(0.5, 1.5)
            n.bias.normal_()
    xs = []
    for h, w in sizes:
        x = torch.randn(N, 256, h, w, device="cuda") * 2 + 0.7
        if channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
        xs.append(x.requires_grad_(True))
    assert dense.group_norm_tokens_supported(xs, norms)
    S = sum(h * w for h, w in sizes)
    go = torch.randn(N, S, 256, device="cuda")
    # reference formulation in fp64
    xr = [x.detach().double().requires_grad_(True) for x in xs]
    pr = [(n.weight.detach().double().requires_grad_(True), n.bias.detach().double().requires_grad_(True)) for n in norms]
    ref = torch.cat([torch.nn.functional.group_norm(x, 32, w, b, n.eps).flatten(2).transpose(1, 2)
                     for x, (w, b), n in zip(xr, pr, norms)], 1)
    ref.backward(go.double())
    if fuse:
        for n in norms:
            for p in (n.weight, n.bias):
                p.grad = torch.full_like(p, 0.25)
                p._fuse_grad = True
    out = dense.group_norm_tokens(xs, norms)
    assert out.shape == (N, S, 256) and out.is_contiguous()
    out.backward(go)
    torch.testing.assert_close(out.double(), ref, rtol=1e-4, atol=1e-4)
    for x, r in zip(xs, xr):
        torch.testing.assert_close(x.grad.double(), r.grad, rtol=1e-3, atol=1e-4 * float(r.grad.abs().max()) + 1e-6)
    for n, (w, b) in zip(norms, pr):
        off = 0.25 if fuse else 0.0
        torch.testing.assert_close(n.weight.grad.double() - off, w.grad, rtol=1e-3, atol=1e-4 * float(w.grad.abs().max()))
        torch.testing.assert_close(n.bias.grad.double() - off, b.grad, rtol=1e-3, atol=1e-4 * float(b.grad.abs().max()))
