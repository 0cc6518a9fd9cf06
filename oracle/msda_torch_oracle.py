"""TEST INFRASTRUCTURE - differentiable PyTorch restatement of MSDeformAttn's sampling core
(semantics of /root/reference/models/ops/functions/ms_deform_attn_func.py:47-65: per level a
zero-padded, align_corners=False bilinear `grid_sample`, weighted by the attention weights).

Used only by tests: it stands in for the CUDA kernel when the *host logic* of the module layer is
checked on a CPU-only box against the golden fixtures (tests/conftest.py::msda_cpu_stub).  It is
itself pinned by tests/test_oracle_msda.py against the fixtures generated from the reference."""
import torch
import torch.nn.functional as F


def msda_core(value, spatial_shapes, sampling_locations, attention_weights):
    """value [N,S,M,D], spatial_shapes list/tensor of (H,W), loc [N,Lq,M,L,P,2], attn [N,Lq,M,L,P]
    -> [N, Lq, M*D]"""
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_locations.shape
    shapes = [(int(h), int(w)) for h, w in (spatial_shapes.tolist() if torch.is_tensor(spatial_shapes) else spatial_shapes)]
    grids = 2 * sampling_locations - 1
    start = 0
    sampled = []
    for lvl, (H, W) in enumerate(shapes):
        v = value[:, start:start + H * W].permute(0, 2, 3, 1).reshape(N * M, D, H, W)
        g = grids[:, :, :, lvl].permute(0, 2, 1, 3, 4).reshape(N * M, Lq, P, 2)
        sampled.append(F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False))
        start += H * W
    sampled = torch.stack(sampled, dim=-2).flatten(-2)                       # [N*M, D, Lq, L*P]
    w = attention_weights.permute(0, 2, 1, 3, 4).reshape(N * M, 1, Lq, L * P)
    out = (sampled * w).sum(-1).view(N, M * D, Lq)
    return out.transpose(1, 2).contiguous()


class CPUFunctionStub:
    """Drop-in for rlipv2_b200.ms_deform_attn.MSDeformAttnFunction on CPU-only test runs."""

    @staticmethod
    def apply(value, shapes, level_start, loc, attn, im2col_step):
        return msda_core(value, shapes, loc, attn)
