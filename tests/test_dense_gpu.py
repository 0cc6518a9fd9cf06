"""GPU tests of the tcgen05 TF32 linear (csrc/dense_tf32.cu) through its C ABI.

Checker: an fp64 torch reference.  Tolerance: each TF32 product carries <= 2^-10 relative error (10-bit
mantissa operands), accumulation is fp32, so |y - y_ref| <= 1.5e-3 * (|x| . |W|^T + |b|) element-wise
(north_star: 1e-3 rel fp32 *with* TF32 tensor-core products being the reference's own arithmetic)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(x, w, b, act):
    y = x.double() @ w.double().t()
    if b is not None:
        y = y + b.double()
    if act == 1:
        y = torch.relu(y)
    elif act == 2:
        y = torch.nn.functional.gelu(y)
    bound = x.double().abs() @ w.double().abs().t() + (b.double().abs() if b is not None else 0)
    return y, bound


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (546, 2048, 256), (512, 768, 3072), (512, 3072, 768),
                                   (1000, 256, 2048), (44446, 2048, 256), (37, 128, 256), (300, 256, 512)])
@pytest.mark.parametrize("act", [0, 1, 2])
def test_linear_tf32_matches_fp64(M, N, K, act):
    from rlipv2_b200 import dense_abi
    g = torch.Generator(device="cuda").manual_seed(M + N + K + act)
    x = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) * K ** -0.5
    b = torch.randn(N, device="cuda", generator=g)
    assert dense_abi.supported(M, N, K)
    y = dense_abi.linear_tf32(x, w, b, act)
    ref, bound = _ref(x, w, b, act)
    err = (y.double() - ref).abs()
    assert torch.isfinite(y).all()
    assert bool((err <= 1.5e-3 * bound + 1e-5).all()), f"max err {float(err.max())}, max ratio {float((err / (bound + 1e-9)).max())}"
    y2 = dense_abi.linear_tf32(x, w, None, 0)
    ref2, bound2 = _ref(x, w, None, 0)
    assert bool(((y2.double() - ref2).abs() <= 1.5e-3 * bound2 + 1e-5).all())


@pytest.mark.parametrize("M,N,K", [(546, 2048, 256), (512, 768, 3072), (300, 256, 512), (37, 128, 256), (1280, 768, 768),
                                   (600, 256, 2048)])
@pytest.mark.parametrize("act", [0, 1, 2])
def test_linear_small_grid_modes_are_bit_identical(M, N, K, act):
    """the small-grid tile / pipeline choices (rlipv2_dense_set_small_mode: 3-stage 128x128, 6-stage 128x128, 8-stage
    128x64) issue the same MMAs in the same K order for every output element: results must be bit-identical"""
    from rlipv2_b200 import dense_abi
    g = torch.Generator(device="cuda").manual_seed(7 * M + N + K + act)
    x = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) * K ** -0.5
    b = torch.randn(N, device="cuda", generator=g)
    keep = dense_abi.small_mode()
    try:
        ys = []
        for mode in (0, 1, 2):
            dense_abi.set_small_mode(mode)
            assert dense_abi.small_mode() == mode
            ys.append(dense_abi.linear_tf32(x, w, b, act))
    finally:
        dense_abi.set_small_mode(keep)
    ref, bound = _ref(x, w, b, act)
    assert bool(((ys[0].double() - ref).abs() <= 1.5e-3 * bound + 1e-5).all())
    assert torch.equal(ys[0], ys[1]) and torch.equal(ys[0], ys[2])


def test_unsupported_shapes_are_reported_not_silently_rerouted():
    from rlipv2_b200 import dense_abi
    assert not dense_abi.supported(64, 100, 256)
    assert not dense_abi.supported(64, 128, 40)
    x = torch.randn(64, 256, device="cuda")
    w = torch.randn(100, 256, device="cuda")
    with pytest.raises(RuntimeError, match="shape not supported"):
        dense_abi.linear_tf32(x, w, None, 0)


def test_autograd_through_dense_seam_tf32():
    from rlipv2_b200 import dense
    dense.set_matmul_precision("tf32")
    try:
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.randn(3, 50, 256, device="cuda", generator=g, requires_grad=True)
        w = (torch.randn(2048, 256, device="cuda", generator=g) / 16).requires_grad_(True)
        b = torch.randn(2048, device="cuda", generator=g, requires_grad=True)
        n0 = __import__("rlipv2_b200.dense_abi", fromlist=["x"]).launch_count()
        y = dense.linear_relu(x, w, b)
        assert __import__("rlipv2_b200.dense_abi", fromlist=["x"]).launch_count() == n0 + 1
        y.square().sum().backward()
        xr, wr, br = (t.detach().double().requires_grad_(True) for t in (x, w, b))
        yr = torch.relu(xr @ wr.t() + br)
        yr.square().sum().backward()
        for a, r in ((x.grad, xr.grad), (w.grad, wr.grad), (b.grad, br.grad)):
            assert float((a.double() - r).norm() / r.norm()) < 3e-3
    finally:
        dense.set_matmul_precision("fp32")


def test_full_step_tf32_close_to_fp32_golden():
    """End-to-end sanity in the benchmark's arithmetic: TF32 tensor-core products (tcgen05 linears +
    cuBLAS TF32) against the fp32 golden fixture, looser tolerance."""
    import numpy as np
    from rlipv2_b200 import dense
    from tests import test_parseda_model as t
    dense.set_matmul_precision("tf32")
    try:
        g, model, criterion, cache, out, loss_dict, total, targets, _ = t._run_step("cuda")
    finally:
        dense.set_matmul_precision("fp32")
    for k in ("pred_obj_logits", "pred_verb_logits", "pred_sub_boxes", "pred_obj_boxes"):
        np.testing.assert_allclose(out[k].detach().cpu().numpy(), g["out_" + k], rtol=3e-2, atol=3e-2)
    np.testing.assert_allclose(float(total), float(g["total_loss"]), rtol=2e-2)


def _bound(a, b):
    """|x|.|w| products bound for TF32 rounding: 2 operands x 2^-11 relative each, fp32 accumulation"""
    return a.abs().double() @ b.abs().double()


@pytest.mark.parametrize("T,N,K,splits", [
    (300, 256, 256, 1), (300, 256, 256, 3), (4001, 128, 256, None), (4001, 2048, 256, None), (4001, 256, 2048, None),
    (1000, 768, 768, 4), (517, 3072, 768, None), (45, 132, 260, 1),
])
def test_wgrad_tf32(T, N, K, splits):
    from rlipv2_b200 import dense_abi
    g = torch.Generator(device="cuda").manual_seed(T + N)
    gy = torch.randn(T, N, device="cuda", generator=g)
    x = torch.randn(T, K, device="cuda", generator=g)
    dw = dense_abi.wgrad_tf32(gy, x, splits)
    ref = gy.double().t() @ x.double()
    err = (dw.double() - ref).abs()
    assert float((err / (1.5e-3 * _bound(gy.t(), x) + 1e-6)).max()) <= 1.0
    assert float(err.max() / ref.abs().max()) < 2e-3


@pytest.mark.parametrize("T,N,K,mask", [
    (300, 256, 2048, True), (300, 2048, 256, False), (4001, 256, 2048, True), (4001, 128, 256, False),
    (517, 768, 3072, False), (45, 132, 260, True), (128, 32, 128, False),
])
def test_dgrad_tf32(T, N, K, mask):
    from rlipv2_b200 import dense_abi
    g = torch.Generator(device="cuda").manual_seed(T + K)
    gy = torch.randn(T, N, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) * N ** -0.5
    h = torch.relu(torch.randn(T, K, device="cuda", generator=g)) if mask else None
    dx, colsum = dense_abi.dgrad_tf32(gy, w, h)
    ref = gy.double() @ w.double()
    bound = 1.5e-3 * _bound(gy, w) + 1e-6
    if mask:
        ref = ref * (h > 0)
        assert float(((colsum.double() - ref.sum(0)).abs() / (bound.sum(0) + 1e-6)).max()) <= 1.0
        assert bool((dx[h <= 0] == 0).all())
    else:
        assert colsum is None
    assert float(((dx.double() - ref).abs() / bound).max()) <= 1.0


def test_linear_rowmask_and_small_wgrad_through_the_seam():
    """dense.linear(..., row_mask) in tf32 mode: masked_fill folded into the GEMM epilogue, its backward (row mask +
    bias gradient in one pass), and the split-K tcgen05 weight gradient (T >= 4096 rows)"""
    from rlipv2_b200 import dense
    try:
        dense.set_matmul_precision("tf32")
        g = torch.Generator(device="cuda").manual_seed(5)
        T = 5000
        x = torch.randn(2, T // 2, 256, device="cuda", generator=g, requires_grad=True)
        w = (torch.randn(256, 256, device="cuda", generator=g) / 16).requires_grad_(True)
        b = torch.randn(256, device="cuda", generator=g).requires_grad_(True)
        mask = torch.rand(2, T // 2, device="cuda", generator=g) < 0.2
        go = torch.randn(2, T // 2, 256, device="cuda", generator=g)
        y = dense.linear(x, w, b, row_mask=mask)
        y.backward(go)
        xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w, b))
        yr = torch.nn.functional.linear(xd, wd, bd).masked_fill(mask[..., None], 0.0)
        yr.backward(go.double())
        assert bool((y[mask] == 0).all())
        rel = lambda a, r: float((a.double() - r).abs().max() / r.abs().max())
        assert rel(y, yr) < 2e-3 and rel(x.grad, xd.grad) < 2e-3 and rel(w.grad, wd.grad) < 2e-3
        assert rel(b.grad, bd.grad) < 1e-5
        assert bool((x.grad[mask] == 0).all())
    finally:
        dense.set_matmul_precision("fp32")


@pytest.mark.parametrize("mode,fuse", [("own", False), ("hybrid", False), ("hybrid", True)])
def test_ffn_relu_fused_backward(monkeypatch, mode, fuse):
    """dense.ffn_relu with the tcgen05 backward (dgrad with fused ReLU mask + bias gradient, split-K wgrad); 'hybrid'
    keeps the plain input gradient on cuBLAS; `fuse`: parameter gradients added into pre-assigned .grad views"""
    from rlipv2_b200 import dense
    try:
        dense.set_matmul_precision("tf32")
        monkeypatch.setattr(dense, "_OWN_BWD", mode == "own")
        monkeypatch.setattr(dense, "_FFN_BWD", "hybrid" if mode == "hybrid" else "cublas")
        g = torch.Generator(device="cuda").manual_seed(6)
        T = 4500
        x = torch.randn(T, 256, device="cuda", generator=g, requires_grad=True)
        w1 = (torch.randn(2048, 256, device="cuda", generator=g) / 16).requires_grad_(True)
        b1 = torch.randn(2048, device="cuda", generator=g).requires_grad_(True)
        w2 = (torch.randn(256, 2048, device="cuda", generator=g) / 45).requires_grad_(True)
        b2 = torch.randn(256, device="cuda", generator=g).requires_grad_(True)
        go = torch.randn(T, 256, device="cuda", generator=g)
        if fuse:
            for t in (w1, b1, w2, b2):
                t.grad = torch.full_like(t, 0.5)                 # "+=": starts non-zero
                t._fuse_grad = True
        y = dense.ffn_relu(x, w1, b1, w2, b2)
        assert y.grad_fn.__class__.__name__ == "_FFNReLUBackward"
        y.backward(go)
        if fuse:
            for t in (w1, b1, w2, b2):
                t.grad = t.grad - 0.5
        ps = [t.detach().double().requires_grad_(True) for t in (x, w1, b1, w2, b2)]
        # the ReLU gate is discontinuous: a hidden unit whose pre-activation is within TF32 rounding of zero may be
        # gated differently in fp64, which moves a gradient entry by a whole term - so the fp64 reference uses the
        # gate the kernel used (h > 0 of the TF32 forward)
        from rlipv2_b200 import dense_abi
        gate = (dense_abi.linear_tf32(x.detach(), w1.detach(), b1.detach(), dense_abi.ACT_RELU) > 0).double()
        yr = torch.nn.functional.linear(torch.nn.functional.linear(ps[0], ps[1], ps[2]) * gate, ps[3], ps[4])
        yr.backward(go.double())
        rel = lambda a, r: float((a.double() - r).abs().max() / r.abs().max())
        assert rel(y, yr) < 3e-3
        for a, r in zip((x, w1, b1, w2, b2), ps):
            assert rel(a.grad, r.grad) < 4e-3
    finally:
        dense.set_matmul_precision("fp32")


@pytest.mark.parametrize("T,mode", [(300, "plain"), (300, "relu"), (5000, "plain"), (5000, "rowmask")])
def test_fused_gradient_accumulation_into_flat_views(T, mode):
    """Parameters marked `_fuse_grad` (the graphed step does that): weight / bias / LayerNorm gradients are ADDED
    straight into the pre-assigned `.grad` views (cuBLAS beta = 1 or the split-K kernel without its zero-fill,
    column sums without the memset) and autograd receives None for them - same values as AccumulateGrad."""
    from rlipv2_b200 import dense
    try:
        dense.set_matmul_precision("tf32")
        g = torch.Generator(device="cuda").manual_seed(7)
        x = torch.randn(T, 256, device="cuda", generator=g, requires_grad=True)
        go = torch.randn(T, 256, device="cuda", generator=g)
        mask = (torch.rand(T, device="cuda", generator=g) < 0.2) if mode == "rowmask" else None

        def params(fuse):
            gg = torch.Generator(device="cuda").manual_seed(8)
            flat = torch.ones(256 * 256 + 256 + 512, device="cuda")          # non-zero start: checks the "+="
            w = torch.nn.Parameter(torch.randn(256, 256, device="cuda", generator=gg) / 16)
            b = torch.nn.Parameter(torch.randn(256, device="cuda", generator=gg))
            lw = torch.nn.Parameter(torch.rand(256, device="cuda", generator=gg) + 0.5)
            lb = torch.nn.Parameter(torch.randn(256, device="cuda", generator=gg))
            o = 0
            for p in (w, b, lw, lb):
                p.grad = flat[o:o + p.numel()].view_as(p)
                o += p.numel()
                if fuse:
                    p._fuse_grad = True
            return flat, w, b, lw, lb

        def run(fuse):
            flat, w, b, lw, lb = params(fuse)
            xx = x.detach().clone().requires_grad_(True)
            if mode == "relu":
                h = dense.linear_relu(xx, w, b)
            else:
                h = dense.linear(xx, w, b, row_mask=mask)
            y = dense.add_layer_norm(h, xx, lw, lb, 1e-5)
            # the weight is used twice: both uses must land in the same view
            y = y + dense.linear(xx, w, b, row_mask=mask) if mode != "relu" else y + dense.linear_relu(xx, w, b)
            y.backward(go)
            dense.join_param_grad_stream()            # small problems form their parameter gradients on a side stream
            return flat, xx.grad

        flat_f, gx_f = run(True)
        flat_r, gx_r = run(False)
        rel = lambda a, r: float((a - r).abs().max() / r.abs().max())
        assert rel(gx_f, gx_r) < 1e-6
        nw = 256 * 256
        assert rel(flat_f[:nw], flat_r[:nw]) < 2e-3                         # weight: different GEMM kernels (TF32)
        assert rel(flat_f[nw:], flat_r[nw:]) < 1e-5                         # bias, LayerNorm gamma / beta
        assert float((flat_r - 1).abs().max()) > 1                          # the gradients are not trivially zero
    finally:
        dense.set_matmul_precision("fp32")


@pytest.mark.parametrize("M,N,K,act,mask", [(44446, 2048, 256, 1, False), (44446, 256, 2048, 0, False), (3001, 128, 256, 0, True),
                                            (20000, 384, 256, 2, False), (777, 1024, 64, 0, False)])
def test_persistent_kernel_is_bit_identical(M, N, K, act, mask):
    """the persistent tcgen05 linear (one CTA per SM walking the tiles, double-buffered TMEM accumulator) issues the same MMAs
    per tile as the one-tile-per-CTA kernel: identical bits, any tile count (more / fewer tiles than SMs, ragged last rows)"""
    from rlipv2_b200 import dense_abi
    g = torch.Generator(device="cuda").manual_seed(M + N)
    x = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) * K ** -0.5
    b = torch.randn(N, device="cuda", generator=g)
    rm = (torch.rand(M, device="cuda", generator=g) < 0.1) if mask else None
    keep = dense_abi.persistent_min_tiles()
    try:
        dense_abi.set_persistent_min_tiles(0)
        ref = dense_abi.linear_tf32(x, w, b, act, rm)
        dense_abi.set_persistent_min_tiles(1)
        assert dense_abi.persistent_min_tiles() == 1
        out = dense_abi.linear_tf32(x, w, b, act, rm)
        out2 = dense_abi.linear_tf32(x, w, b, act, rm)
    finally:
        dense_abi.set_persistent_min_tiles(keep)
    assert torch.equal(out, ref) and torch.equal(out2, ref)
    r64, bound = _ref(x, w, b, act)
    if rm is not None:
        r64 = r64.masked_fill(rm[:, None], 0.0)
    assert bool(((out.double() - r64).abs() <= 1.5e-3 * bound + 1e-5).all())


@pytest.mark.parametrize("T,N,K", [(44446, 256, 2048), (5000, 96, 512), (1000, 256, 256)])
def test_persistent_gated_dgrad_is_bit_identical(T, N, K):
    """dgrad_mask_persistent_kernel (128 x 256 tiles, one CTA per SM) against the one-tile-per-CTA EPI_MASK kernel: the gated
    input gradient and the per-row-tile column sums, bit for bit; and both against fp64"""
    from rlipv2_b200 import dense_abi
    g = torch.Generator(device="cuda").manual_seed(T + K)
    gy = torch.randn(T, N, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) * N ** -0.5
    h = torch.relu(torch.randn(T, K, device="cuda", generator=g))
    keep = dense_abi.persistent_min_tiles()
    try:
        dense_abi.set_persistent_min_tiles(0)
        dx0, cs0 = dense_abi.dgrad_tf32(gy, w, h)
        dense_abi.set_persistent_min_tiles(1)
        dense_abi.set_persistent_dgrad(True)
        dx1, cs1 = dense_abi.dgrad_tf32(gy, w, h)
        dx2, cs2 = dense_abi.dgrad_tf32(gy, w, h)
    finally:
        dense_abi.set_persistent_min_tiles(keep)
        dense_abi.set_persistent_dgrad(False)
    assert torch.equal(dx1, dx0) and torch.equal(dx2, dx0)
    # the per-tile column sums are identical; their sum over the row tiles is formed with fp32 atomics (order varies)
    torch.testing.assert_close(cs1, cs0, rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(cs2, cs0, rtol=1e-5, atol=1e-4)
    ref = (gy.double() @ w.double()) * (h > 0)
    bound = 1.5e-3 * _bound(gy, w) + 1e-6
    assert float(((dx1.double() - ref).abs() / bound).max()) <= 1.0
    assert float(((cs1.double() - ref.sum(0)).abs() / (bound.sum(0) + 1e-6)).max()) <= 1.0
