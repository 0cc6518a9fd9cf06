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
    assert dense._GroupNormTokensMulti is not None
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


def _sa_inputs(B, H, T, seed, pad=True):
    g = torch.Generator(device="cuda").manual_seed(seed)
    mk = lambda: torch.randn(B, T, H * 64, device="cuda", generator=g).view(B, T, H, 64).transpose(1, 2)   # HF's views
    q, k, v = mk(), mk(), mk()
    lens = torch.randint(1, T + 1, (B,), device="cuda", generator=g) if pad else torch.full((B,), T, device="cuda")
    keep = (torch.arange(T, device="cuda")[None] < lens[:, None]).float()
    mask = ((1.0 - keep) * torch.finfo(torch.float32).min)[:, None, None, :]
    return q, k, v, mask


@pytest.mark.parametrize("B,H,T", [(256, 12, 5), (7, 3, 8), (1, 1, 1), (33, 12, 3)])
def test_short_attention_matches_hf_eager(B, H, T):
    """rlipv2_short_attention_* (text tower, label strings) against HF's eager_attention_forward, forward + backward"""
    from transformers.models.roberta.modeling_roberta import eager_attention_forward
    from rlipv2_b200.text_encoder import _short_attention_forward

    M = lambda: torch.nn.Module().eval()                 # eval mode: the dropout argument must be ignored
    torch.manual_seed(B * 31 + T)
    q, k, v, mask = _sa_inputs(B, H, T, seed=B + T)
    go = torch.randn(B, T, H, 64, device="cuda")
    res = []
    for fn in (eager_attention_forward, _short_attention_forward):
        qq, kk, vv = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
        out, _ = fn(M(), qq, kk, vv, mask, dropout=0.1, scaling=0.125)
        assert out.shape == (B, T, H, 64)
        (out * go).sum().backward()
        res.append((out.detach(), qq.grad, kk.grad, vv.grad))
    for a, b in zip(res[1], res[0]):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5 * float(b.abs().max()) + 1e-6)


def test_short_attention_dropout_is_consistent_between_forward_and_backward():
    from rlipv2_b200.text_encoder import _ShortAttention
    torch.manual_seed(1234)
    B, H, T, p = 64, 12, 5, 0.1
    q, k, v, mask = _sa_inputs(B, H, T, seed=1, pad=False)
    seed = torch.tensor([12345], dtype=torch.int64, device="cuda")
    m2 = mask.reshape(B, T).contiguous()
    run = lambda qq, kk, vv, salt=3: _ShortAttention.apply(qq, kk, vv, m2, 0.125, p, seed, salt)
    # keep rate: uniform probabilities (q = 0), v = 1 -> every output element = (#kept / T) / (1 - p), mean 1
    ones = torch.ones_like(v)
    o = run(torch.zeros_like(q), k, ones)
    assert abs(float(o.mean()) - 1.0) < 0.02
    frac_dropped = float((run(torch.zeros_like(q), k, ones, salt=4) != o).float().mean())
    assert frac_dropped > 0.1                                        # another call site draws another mask
    # the mask is a function of (seed, salt, element): linear in v for fixed q, k
    v2 = torch.randn_like(v)
    torch.testing.assert_close(run(q, k, v + v2), run(q, k, v) + run(q, k, v2), rtol=1e-4, atol=1e-4)
    # adjoint identity through the dropped attention: <go, out(v)> = <dv, v>, and a directional derivative in q, k
    qq, kk, vv = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
    go = torch.randn(B, T, H, 64, device="cuda")
    out = run(qq, kk, vv)
    (out * go).sum().backward()
    lhs, rhs = float((go.double() * out.double()).sum()), float((vv.grad.double() * v.double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), 1.0)
    dq, dk = torch.randn_like(q), torch.randn_like(k)
    eps = 1e-2
    f = lambda s: float((run(q + s * eps * dq, k + s * eps * dk, v).double() * go.double()).sum())
    num = (f(1) - f(-1)) / (2 * eps)
    ana = float((qq.grad.double() * dq.double()).sum() + (kk.grad.double() * dk.double()).sum())
    # |ana| is ~ N(0, 300) for these shapes (a wrong mask / factor moves it by O(100)); fp32 evaluation of f leaves ~0.05
    assert abs(num - ana) <= 2e-2 * max(abs(ana), 50.0), (num, ana)
