"""GPU: split-K tcgen05 forward linear (rlipv2_dense_linear_splitk_tf32) on the small-M / long-K shapes of ALIF and the
RobertaLayer against fp64, at the TF32 tolerance of the plain kernel's test, and through dense.linear with
RLIPV2_SPLITK_FWD.  Runs last: written after round 1's GPU budget was spent."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(546, 256, 2048), (512, 768, 2048), (512, 2048, 768), (512, 768, 3072), (300, 260, 1024),
                                   (1, 128, 768)])
@pytest.mark.parametrize("bias", [True, False])
def test_splitk_linear_matches_fp64(M, N, K, bias):
    from rlipv2_b200 import dense_abi
    torch.manual_seed(M + N + K)
    x = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") * K ** -0.5
    b = torch.randn(N, device="cuda") if bias else None
    want = x.double() @ w.double().t() + (b.double() if bias else 0)
    for splits in (2, dense_abi.splitk_splits(M, N, K), K // 32):
        if splits < 1:
            continue
        y = dense_abi.linear_splitk_tf32(x, w, b, splits)
        err = (y.double() - want).abs().max().item()
        assert err < 2e-3 * want.abs().max().item(), (splits, err)
    plain = dense_abi.linear_tf32(x, w, b, 0) if N % 128 == 0 else None
    if plain is not None:                                # same products, different summation order over K slices
        assert (plain - y).abs().max().item() < 1e-3 * want.abs().max().item()


def test_dense_linear_routes_long_k_through_splitk(monkeypatch):
    from rlipv2_b200 import dense, dense_abi
    try:
        dense.set_matmul_precision("tf32")
        monkeypatch.setattr(dense, "_SPLITK_FWD", True)
        x = torch.randn(2, 256, 2048, device="cuda", requires_grad=True)
        lin = torch.nn.Linear(2048, 768).cuda()
        n0 = dense_abi.launch_count()
        y = dense.linear(x, lin.weight, lin.bias)
        assert dense_abi.launch_count() == n0 + 1
        want = torch.nn.functional.linear(x.double(), lin.weight.double(), lin.bias.double())
        assert (y.double() - want).abs().max().item() < 2e-3 * want.abs().max().item()
        y.sum().backward()                               # backward path unchanged
        assert torch.isfinite(x.grad).all() and lin.weight.grad is not None
    finally:
        dense.set_matmul_precision("fp32")
