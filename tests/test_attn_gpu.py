"""GPU parity of the fused tcgen05 attention cores (csrc/attn_tf32.cu, include/rlipv2_attn.h) through the C ABI.

Checker: the reference's own op chain (bmm -> softmax -> dropout -> bmm; fuse_helper.py:395-445, modeling_roberta.py:185-241,
nn.MultiheadAttention) evaluated by torch in fp64 on the same inputs.  Tolerance: the products are TF32 (2^-11 relative per
operand, fp32 accumulation), so results are held to a few 1e-3 of the output scale - the bound the tcgen05 linear tests use."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

# (B, H, Tq, Nk, D, key bias?)  - the shapes of the step at BASELINE config 2 and some ragged ones
CASES = [
    (2, 8, 273, 256, 256, False),      # ALIF vision -> language (13 x 21 image tokens, 256 labels)
    (2, 8, 256, 273, 256, False),      # ALIF language -> vision: 273 keys -> two TMEM output slices
    (2, 12, 256, 256, 64, True),       # RobertaLayer, additive -10000 key mask
    (2, 8, 300, 300, 32, False),       # pair decoder query self-attention
    (2, 8, 150, 150, 32, False),       # verb decoder
    (1, 2, 37, 61, 64, True),          # ragged: one partial query tile, keys not a multiple of 32
    (3, 1, 129, 33, 128, False),
]


def _inputs(B, H, Tq, Nk, D, bias, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    q = torch.randn(B, Tq, H * D, device="cuda", generator=g)
    k = torch.randn(B, Nk, H * D, device="cuda", generator=g)
    v = torch.randn(B, Nk, H * D, device="cuda", generator=g)
    kb = None
    if bias:
        kb = torch.zeros(B, Nk, device="cuda")
        kb[:, Nk - Nk // 5:] = -10000.0
        kb[0, 1] = -10000.0
    return q, k, v, kb


def _ref(q, k, v, kb, H, scale, mask=None, inv_keep=1.0):
    """fp64 reference; `mask` [B, H, Tq, Nk] bool = kept elements of the dropout"""
    B, Tq, C = q.shape
    D = C // H
    qh, kh, vh = (t.double().reshape(B, -1, H, D).transpose(1, 2) for t in (q, k, v))
    s = torch.matmul(qh, kh.transpose(-1, -2)) * scale
    if kb is not None:
        s = s + kb.double()[:, None, None, :]
    p = torch.softmax(s, -1)
    if mask is not None:
        p = p * mask.double() * inv_keep
    return torch.matmul(p, vh).transpose(1, 2).reshape(B, Tq, C)


@pytest.mark.parametrize("B,H,Tq,Nk,D,bias", CASES)
def test_forward_matches_fp64(B, H, Tq, Nk, D, bias):
    from rlipv2_b200 import attn_abi
    q, k, v, kb = _inputs(B, H, Tq, Nk, D, bias)
    scale = D ** -0.5
    n0 = attn_abi.launch_count()
    out, stats, _ = attn_abi.forward(q, k, v, H, kb, scale, 0.0, None, 0)
    assert attn_abi.launch_count() == n0 + 1
    ref = _ref(q, k, v, kb, H, scale)
    err = float((out.double() - ref).abs().max())
    assert err <= 4e-3 * float(ref.abs().max()), err
    # row statistics the backward uses: max and sum of exp of the scaled, biased scores
    qh, kh = (t.double().reshape(B, -1, H, D).transpose(1, 2) for t in (q, k))
    s = torch.matmul(qh, kh.transpose(-1, -2)) * scale
    if kb is not None:
        s = s + kb.double()[:, None, None, :]
    m = s.max(-1).values.reshape(B * H, Tq)
    torch.testing.assert_close(stats[..., 0].double(), m, rtol=0, atol=2e-2)
    lse = (stats[..., 0].double() + stats[..., 1].double().log())
    torch.testing.assert_close(lse, torch.logsumexp(s, -1).reshape(B * H, Tq), rtol=0, atol=2e-2)


@pytest.mark.parametrize("B,H,Tq,Nk,D,bias", CASES)
def test_backward_matches_fp64_autograd(B, H, Tq, Nk, D, bias):
    from rlipv2_b200 import attn_abi
    q, k, v, kb = _inputs(B, H, Tq, Nk, D, bias, seed=1)
    scale = D ** -0.5
    g = torch.randn(B, Tq, H * D, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    out, stats, seed_used = attn_abi.forward(q, k, v, H, kb, scale, 0.0, None, 0)
    dq, dk, dv = attn_abi.backward(q, k, v, H, kb, out, g, stats, scale, 0.0, seed_used, 0)
    qr, kr, vr = (t.detach().double().requires_grad_(True) for t in (q, k, v))
    _ref(qr, kr, vr, kb, H, scale).backward(g.double())
    for name, a, r in (("dq", dq, qr.grad), ("dk", dk, kr.grad), ("dv", dv, vr.grad)):
        rel = float((a.double() - r).norm() / r.norm())
        assert rel < 4e-3, (name, rel)
        assert float((a.double() - r).abs().max()) <= 2e-2 * float(r.abs().max()), name
    # accumulate mode: the same gradients added onto pre-filled buffers (ALIF's two directions share dq / dk)
    base = [torch.full_like(t, 0.5) for t in (dq, dk, dv)]
    attn_abi.backward(q, k, v, H, kb, out, g, stats, scale, 0.0, seed_used, 0, *base)
    for a, b_ in zip((dq, dk, dv), base):
        torch.testing.assert_close(b_, a + 0.5, rtol=1e-5, atol=1e-5)


def test_operands_are_read_in_place_from_a_fused_projection():
    """the decoder hands q / k as column slices of one [B, T, 2C] projection (row stride 2C): no copies"""
    from rlipv2_b200 import attn_abi
    B, H, T, D = 2, 8, 150, 32
    C = H * D
    g = torch.Generator(device="cuda").manual_seed(2)
    qk = torch.randn(B, T, 2 * C, device="cuda", generator=g)
    v = torch.randn(B, T, C, device="cuda", generator=g)
    q, k = qk[..., :C], qk[..., C:]
    assert attn_abi.usable(q) and attn_abi.usable(k) and not k.is_contiguous()
    out, stats, su = attn_abi.forward(q, k, v, H, None, D ** -0.5, 0.0, None, 0)
    ref = _ref(q, k, v, None, H, D ** -0.5)
    assert float((out.double() - ref).abs().max()) <= 4e-3 * float(ref.abs().max())
    go = torch.randn_like(out)
    dq, dk, dv = attn_abi.backward(q, k, v, H, None, out, go, stats, D ** -0.5, 0.0, su, 0)
    qr, kr, vr = (t.detach().double().requires_grad_(True) for t in (q, k, v))
    _ref(qr, kr, vr, None, H, D ** -0.5).backward(go.double())
    for a, r in ((dq, qr.grad), (dk, kr.grad), (dv, vr.grad)):
        assert float((a.double() - r).norm() / r.norm()) < 4e-3


def test_dropout_mask_is_hashed_and_regenerated_by_the_backward():
    """V = identity makes the output the dropped probability map itself: the kept fraction is 1 - p, kept entries equal
    softmax / (1 - p), another seed gives another mask, and the backward reproduces autograd under that exact mask."""
    from rlipv2_b200 import attn_abi
    B, H, Tq, Nk, D = 1, 1, 200, 256, 256
    p = 0.1
    g = torch.Generator(device="cuda").manual_seed(3)
    q = torch.randn(B, Tq, D, device="cuda", generator=g)
    k = torch.randn(B, Nk, D, device="cuda", generator=g)
    v = torch.eye(Nk, D, device="cuda")[None].contiguous()
    scale = D ** -0.5
    seed = torch.tensor([1234], dtype=torch.int64, device="cuda")
    out, stats, su = attn_abi.forward(q, k, v, H, None, scale, p, seed, 7)
    assert int(su) == 1234
    probs = torch.softmax((q.double() @ k.double().transpose(1, 2)) * scale, -1)
    mask = out[0] != 0
    assert abs(float(mask.double().mean()) - (1 - p)) < 0.01
    kept = out[0].double()[mask]
    assert float((kept - probs[0][mask] / (1 - p)).abs().max()) <= 4e-3 * float(probs.max()) / (1 - p)
    out_b, _, _ = attn_abi.forward(q, k, v, H, None, scale, p, seed, 7)
    assert torch.equal(out, out_b)                                               # same seed and salt: same mask
    out2, _, _ = attn_abi.forward(q, k, v, H, None, scale, p, seed + 1, 7)
    out3, _, _ = attn_abi.forward(q, k, v, H, None, scale, p, seed, 8)
    assert not torch.equal(out2 != 0, mask) and not torch.equal(out3 != 0, mask)
    # backward under the regenerated mask, with a general V
    v2 = torch.randn(B, Nk, D, device="cuda", generator=g)
    out4, stats4, su4 = attn_abi.forward(q, k, v2, H, None, scale, p, seed, 7)
    go = torch.randn_like(out4)
    dq, dk, dv = attn_abi.backward(q, k, v2, H, None, out4, go, stats4, scale, p, su4, 7)
    qr, kr, vr = (t.detach().double().requires_grad_(True) for t in (q, k, v2))
    ref = _ref(qr, kr, vr, None, H, scale, mask=mask[None, None], inv_keep=1 / (1 - p))
    assert float((out4.double() - ref).abs().max()) <= 4e-3 * float(ref.abs().max())
    ref.backward(go.double())
    for name, a, r in (("dq", dq, qr.grad), ("dk", dk, kr.grad), ("dv", dv, vr.grad)):
        assert float((a.double() - r).norm() / r.norm()) < 4e-3, name


def test_seam_routes_tf32_mode_to_the_kernels_and_matches_the_torch_path():
    from rlipv2_b200 import attn_abi, dense
    try:
        B, H, Tv, Tl, D = 2, 8, 273, 256, 256
        g = torch.Generator(device="cuda").manual_seed(4)
        mk = lambda t: torch.randn(B, t, H * D, device="cuda", generator=g)
        q, k, vv, vl = mk(Tv), mk(Tl), mk(Tv), mk(Tl)
        go_v, go_l = mk(Tv), mk(Tl)
        res = {}
        for mode in ("fp32", "tf32"):
            dense.set_matmul_precision(mode)
            leaves = [t.clone().requires_grad_(True) for t in (q, k, vv, vl)]
            n0 = attn_abi.launch_count()
            ov, ol = dense.bi_attention(*leaves, H, D ** -0.5, 0.1, False)
            (ov * go_v).sum().backward(retain_graph=True)
            (ol * go_l).sum().backward()
            res[mode] = [ov, ol] + [t.grad for t in leaves]
            launched = attn_abi.launch_count() - n0
            assert launched == (0 if mode == "fp32" else 2 + 2 * 2), (mode, launched)   # fwd; dS + one 3-problem GEMM launch
        for a, r in zip(res["tf32"], res["fp32"]):
            assert float((a - r).norm() / r.norm()) < 4e-3
    finally:
        dense.set_matmul_precision("fp32")


def test_unsupported_problem_is_reported():
    from rlipv2_b200 import attn_abi
    assert not attn_abi.supported(1, 1, 64, 600, 64) and not attn_abi.supported(1, 1, 64, 64, 48)
    q = torch.randn(1, 64, 48, device="cuda")
    with pytest.raises(RuntimeError, match="shape not supported"):
        attn_abi.forward(q, q, q, 1, None, 1.0, 0.0, None, 0)


def test_modules_in_tf32_mode_match_the_reference_goldens():
    """ALIF block and RobertaLayer against the fixtures the reference's own modules produced
    (tests/golden/parseda_alif.npz, parseda_roberta.npz), in the benchmark's arithmetic: every contraction of both blocks
    on the tcgen05 kernels (projections + fused attention cores)."""
    import numpy as np
    from oracle.detfill import det_fill_
    from rlipv2_b200 import attn_abi, dense
    from rlipv2_b200.alif import RLIPv2_VLFuse
    from rlipv2_b200.roberta_layer import RobertaLayer
    from rlipv2_b200.text_encoder import roberta_base_config
    from tests.test_parseda_model import _args, _load
    try:
        dense.set_matmul_precision("tf32")
        g = _load("parseda_alif.npz")
        fuse = det_fill_(RLIPv2_VLFuse(_args("cuda")), seed=1).eval().to("cuda")
        t = lambda k: torch.from_numpy(g[k]).to("cuda")
        n0 = attn_abi.launch_count()
        with torch.no_grad():
            out = fuse({"visual": {"src": t("v"), "padding_mask": t("mask_v"), "pos": t("pos")},
                        "lang": {"hidden": t("l"), "masks": t("mask_l")}})
        assert attn_abi.launch_count() == n0 + 2
        for a, key in ((out["visual"]["src"], "out_v"), (out["lang"]["hidden"], "out_l")):
            ref = g[key]
            assert float(np.abs(a.cpu().numpy() - ref).max()) <= 5e-3 * float(np.abs(ref).max()), key
        g = _load("parseda_roberta.npz")
        layer = det_fill_(RobertaLayer(roberta_base_config()), seed=2).eval().to("cuda")
        n0 = attn_abi.launch_count()
        with torch.no_grad():
            y = layer(torch.from_numpy(g["x"]).to("cuda"), attention_mask=torch.from_numpy(g["mask"]).to("cuda"))
        assert attn_abi.launch_count() == n0 + 1
        assert float(np.abs(y.cpu().numpy() - g["y"]).max()) <= 5e-3 * float(np.abs(g["y"]).max())
    finally:
        dense.set_matmul_precision("fp32")
