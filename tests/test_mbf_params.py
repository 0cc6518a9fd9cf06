"""MultiBranchFusion (dab_deformable/deformable_transformer.py:1025-1068): the stacked-GEMM forward and the
parameter-gradient paths of `_StackedBranchParams` (per-parameter gradients / two strided adds into the flat gradient
buffer) against the reference's literal per-branch formulation.  CPU (host logic; the arithmetic is torch's)."""
import torch
import torch.nn.functional as F

from rlipv2_b200.parseda_transformer import MultiBranchFusion


def _literal(m, a, b):
    """relu(sum_c fc_3[c](relu(fc_1[c](a) * fc_2[c](b)))) exactly as the reference's list comprehension"""
    return F.relu(torch.stack([f3(F.relu(f1(a) * f2(b))) for f1, f2, f3 in zip(m.fc_1, m.fc_2, m.fc_3)]).sum(0))


def _run(fused):
    torch.manual_seed(0)
    m = MultiBranchFusion(32, 32, 32, 4).double()
    a = torch.randn(2, 7, 32, dtype=torch.double, requires_grad=True)
    b = torch.randn(2, 7, 32, dtype=torch.double, requires_grad=True)
    gout = torch.randn(2, 7, 32, dtype=torch.double)
    ref = _literal(m, a, b)
    ref_grads = torch.autograd.grad(ref, list(m.parameters()) + [a, b], gout)
    if fused:
        # the layout FlatParams gives the step: .grad = adjacent views of one flat buffer, in named_parameters order
        flat_grad = torch.full((sum(p.numel() for p in m.parameters()),), 0.25, dtype=torch.double)   # must accumulate
        o = 0
        for p in m.parameters():
            p.grad = flat_grad[o:o + p.numel()].view_as(p)
            o += p.numel()
            p._fuse_grad = True
        from rlipv2_b200.parseda_transformer import _StackedBranchParams
        for fc in (m.fc_1, m.fc_2, m.fc_3):       # the strided-add path is the one exercised below
            assert _StackedBranchParams._adjacent_grad_views(tuple(t for l in fc for t in (l.weight, l.bias))) is not None
    out = m(a, b)
    out.backward(gout)
    torch.testing.assert_close(out, ref)
    for p, g in zip(list(m.parameters()) + [a, b], ref_grads):
        want = g + 0.25 if (fused and p.dim() and p is not a and p is not b) else g
        torch.testing.assert_close(p.grad, want, msg=lambda s: f"{tuple(p.shape)}: {s}")


def test_mbf_matches_literal_reference_formulation():
    _run(fused=False)


def test_mbf_fused_gradient_accumulation_into_flat_views():
    _run(fused=True)
