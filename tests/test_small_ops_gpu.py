"""Small fused kernels of the DAB decoder and the backward graph's host-flag wait (csrc/fused_ops.cu) against the
torch formulas they replace (/root/reference/util/misc.py:460-464, dab_deformable/deformable_transformer.py:1777-1802,
1511-1541)."""
import time

import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda")


def test_box_refine_matches_torch():
    from rlipv2_b200 import dense
    dev = _dev()
    g = torch.Generator(device="cpu").manual_seed(0)
    delta = (torch.randn(2, 150, 4, generator=g) * 3).to(dev).requires_grad_(True)
    ref = torch.rand(2, 150, 4, generator=g)
    ref.view(-1)[:8] = torch.tensor([0.0, 1.0, -0.2, 1.3, 1e-7, 1 - 1e-7, 0.5, 1e-5])      # clamp branches
    ref = ref.to(dev).requires_grad_(True)
    y = dense.box_refine(delta, ref)
    assert isinstance(y.grad_fn, dense._BoxRefine._backward_cls)
    w = torch.randn_like(y)
    (y * w).sum().backward()
    d2 = delta.detach().clone().requires_grad_(True)
    r2 = ref.detach().clone().requires_grad_(True)
    y2 = (d2 + dense._inverse_sigmoid_torch(r2)).sigmoid()
    (y2 * w).sum().backward()
    torch.testing.assert_close(y, y2, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(delta.grad, d2.grad, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(ref.grad, r2.grad, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("n", [2, 4])
def test_sine_embed_matches_torch(n):
    from rlipv2_b200 import dense
    dev = _dev()
    pos = torch.rand(2, 300, n, device=dev, requires_grad=True)
    e = dense.sine_embed(pos)
    assert e.shape == (2, 300, n * 128) and isinstance(e.grad_fn, dense._SineEmbed._backward_cls)
    p2 = pos.detach().clone().requires_grad_(True)
    e2 = dense._sine_embed_torch(p2)
    # arguments reach 2*pi; sinf/cosf and the IEEE division are the same functions torch calls
    torch.testing.assert_close(e, e2, rtol=0, atol=2e-6)
    w = torch.randn_like(e)
    (e * w).sum().backward()
    (e2 * w).sum().backward()
    torch.testing.assert_close(pos.grad, p2.grad, rtol=1e-5, atol=1e-5)


def test_wait_host_flag_holds_the_stream_until_published():
    from rlipv2_b200 import fused_abi
    dev = _dev()
    flag = torch.zeros(1, dtype=torch.int32).pin_memory()
    seq = torch.zeros(1, dtype=torch.int32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    marker = torch.zeros(1, device=dev)
    host = torch.zeros(1).pin_memory()
    s = torch.cuda.Stream()
    for step in (1, 2):
        with torch.cuda.stream(s):
            fused_abi.wait_host_flag(flag, seq, err, timeout_s=20.0)
            marker.add_(1.0)
            host.copy_(marker, non_blocking=True)
            done = torch.cuda.Event()
            done.record()
        time.sleep(0.05)
        assert not done.query() and host.item() == step - 1          # the stream is parked behind the flag
        flag[0] = step                                                # publish
        done.synchronize()
        assert host.item() == step
    assert seq.item() == 2 and err.item() == 0


def test_wait_host_flag_times_out_instead_of_hanging():
    from rlipv2_b200 import fused_abi
    dev = _dev()
    flag = torch.zeros(1, dtype=torch.int32).pin_memory()
    seq = torch.zeros(1, dtype=torch.int32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    fused_abi.wait_host_flag(flag, seq, err, timeout_s=0.05)
    torch.cuda.synchronize()
    assert err.item() == 1 and seq.item() == 1
