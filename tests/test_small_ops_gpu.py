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


@pytest.mark.parametrize("case", ["random", "disjoint", "nested", "identical"])
def test_box_pair_loss_matches_torch_autograd(case):
    """rlipv2_box_pair_loss_f32 vs the criterion's torch formulation (hoi.py:4162-4193, util/box_ops.py:19-73): forward
    values bit-identical (same op order, no FMA contraction), gradients against torch.autograd in fp64"""
    from rlipv2_b200.criterion import SetCriterionHOI, _BoxPairLoss
    from rlipv2_b200.nested import box_cxcywh_to_xyxy
    dev = _dev()
    g = torch.Generator(device="cpu").manual_seed(11)
    R = 301
    mk = lambda: torch.cat((torch.rand(R, 2, generator=g) * 0.6 + 0.2, torch.rand(R, 2, generator=g) * 0.5 + 0.02), 1)
    s, t = mk(), mk()
    if case == "disjoint":
        t[:, :2] = s[:, :2] + 0.8
    elif case == "nested":
        t[:, :2] = s[:, :2]
        t[:, 2:] = s[:, 2:] * 0.3
    elif case == "identical":
        t = s.clone()                                     # every min / max is a tie: gradients split evenly
    s_d = s.to(dev).requires_grad_(True)
    l1, gl = _BoxPairLoss.apply(s_d, t.to(dev))
    w1, w2 = torch.randn(R, generator=g), torch.randn(R, generator=g)
    ((l1 * w1.to(dev)).sum() + (gl * w2.to(dev)).sum()).backward()
    # forward: the same fp32 expression evaluated by torch on the device
    s_t = s.to(dev)
    ref_l1 = (s_t - t.to(dev)).abs().sum(-1)
    ref_gl = 1 - SetCriterionHOI._paired_giou(box_cxcywh_to_xyxy(s_t), box_cxcywh_to_xyxy(t.to(dev)))
    torch.testing.assert_close(l1, ref_l1, rtol=1e-6, atol=1e-6)
    assert torch.equal(gl, ref_gl)
    # gradient: autograd through the torch formulation in fp64
    s64 = s.double().requires_grad_(True)
    l1_64 = (s64 - t.double()).abs().sum(-1)
    gl_64 = 1 - SetCriterionHOI._paired_giou(box_cxcywh_to_xyxy(s64), box_cxcywh_to_xyxy(t.double()))
    ((l1_64 * w1.double()).sum() + (gl_64 * w2.double()).sum()).backward()
    torch.testing.assert_close(s_d.grad.cpu().double(), s64.grad, rtol=1e-4, atol=1e-4)
