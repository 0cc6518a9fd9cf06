"""Host-side mirror of the reference's MSDeformAttn operator surface.

  MSDeformAttnFunction  <- /root/reference/models/ops/functions/ms_deform_attn_func.py:25-44
  MSDeformAttn          <- /root/reference/models/ops/modules/ms_deform_attn.py:34-119
                           (identical copy: models/dab_deformable/ops/modules/ms_deform_attn.py)

Same names, argument order, parameter names / shapes / initialisation (so reference checkpoints
load), same error behaviour.  The sampling + aggregation runs in the sm_100a kernels of
csrc/msda.cu through the C ABI (include/rlipv2_msda.h); there is no PyTorch fallback.
"""
import math
import os
import warnings

import torch
import torch.nn.functional as F
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.init import constant_, xavier_uniform_

from . import dense
from .dropin import MultiScaleDeformableAttention as MSDA


class MSDeformAttnFunction(Function):
    """autograd glue, identical contract to ms_deform_attn_func.py:25-44."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        output = MSDA.ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index,
                                             sampling_locations, attention_weights, ctx.im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index,
                              sampling_locations, attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, level_start, sampling_locations, attention_weights = ctx.saved_tensors
        grad_value, grad_sampling_loc, grad_attn_weight = MSDA.ms_deform_attn_backward(
            value, shapes, level_start, sampling_locations, attention_weights,
            grad_output.contiguous(), ctx.im2col_step)
        return grad_value, None, None, grad_sampling_loc, grad_attn_weight, None


class MSDeformAttnProjFunction(Function):
    """Fused-prologue op (include/rlipv2_msda.h `rlipv2_msda_proj_*`): the softmax over the 16 attention
    logits and `loc = ref + offset / (W_l, H_l)` of ms_deform_attn.py:102-109 happen inside the sampling
    kernels, forward and backward, so neither `sampling_locations` nor `attention_weights` (68 MB per
    encoder call at 800x1333, batch 2) is ever materialised.  `reference_points` is not differentiated (the
    encoder's are constants of the image size)."""

    @staticmethod
    def forward(ctx, value, spatial_shapes, level_start_index, reference_points, proj):
        from . import msda_abi
        N, Lq = proj.shape[:2]
        out = torch.empty((N, Lq, value.shape[2] * value.shape[3]), dtype=value.dtype, device=value.device)
        msda_abi.proj_forward(value, spatial_shapes, level_start_index, reference_points, proj, out)
        ctx.save_for_backward(value, spatial_shapes, level_start_index, reference_points, proj)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        from . import msda_abi
        value, shapes, level_start, reference_points, proj = ctx.saved_tensors
        grad_value = torch.empty_like(value)          # zero-filled inside the library call
        grad_proj = torch.empty_like(proj)            # fully written by the kernel
        msda_abi.proj_backward(value, shapes, level_start, reference_points, proj, grad_output.contiguous(),
                               grad_value, grad_proj)
        return grad_value, None, None, None, grad_proj


_FUSED_PROLOGUE = os.environ.get("RLIPV2_MSDA_FUSED_PROLOGUE", "1") != "0"     # A/B switch for measurements
_FUSED_PROLOGUE_REF4 = os.environ.get("RLIPV2_MSDA_FUSED_PROLOGUE_REF4", "1") != "0"   # decoder cross-attention (4-d anchors)


def _is_power_of_2(n):
    if (not isinstance(n, int)) or (n < 0):
        raise ValueError("invalid input for _is_power_of_2: {} (type: {})".format(n, type(n)))
    return (n & (n - 1) == 0) and n != 0


class MSDeformAttn(nn.Module):
    """Multi-scale deformable attention module (ms_deform_attn.py:34-119).

    Parameters (state_dict keys): sampling_offsets.{weight,bias} [M*L*P*2, C], attention_weights
    [M*L*P, C], value_proj [C, C], output_proj [C, C].
    """

    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError("d_model must be divisible by n_heads, but got {} and {}".format(d_model, n_heads))
        if not _is_power_of_2(d_model // n_heads):
            warnings.warn("You'd better set d_model in MSDeformAttn to make the dimension of each "
                          "attention head a power of 2 which is more efficient in our CUDA implementation.")
        self.im2col_step = 64
        self.d_model = d_model
        self.n_levels = n_levels
        self.n_heads = n_heads
        self.n_points = n_points
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self._reset_parameters()

    def _reset_parameters(self):
        # ms_deform_attn.py:66-80: zero offset weights, a ring of directions scaled by the point
        # index as offset bias, zero attention logits, xavier value / output projections.
        constant_(self.sampling_offsets.weight.data, 0.)
        thetas = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        grid_init = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid_init = (grid_init / grid_init.abs().max(-1, keepdim=True)[0]).view(
            self.n_heads, 1, 1, 2).repeat(1, self.n_levels, self.n_points, 1)
        for i in range(self.n_points):
            grid_init[:, :, i, :] *= i + 1
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(grid_init.view(-1))
        constant_(self.attention_weights.weight.data, 0.)
        constant_(self.attention_weights.bias.data, 0.)
        xavier_uniform_(self.value_proj.weight.data)
        constant_(self.value_proj.bias.data, 0.)
        xavier_uniform_(self.output_proj.weight.data)
        constant_(self.output_proj.bias.data, 0.)

    def project_value(self, input_flatten, input_padding_mask=None):
        """value_proj + `masked_fill(padding_mask, 0)` (:98-100): the mask rides in the GEMM epilogue.  Callers whose
        `input_flatten` is ready long before the query (the decoders: the same encoder memory for every layer) can
        evaluate this ahead of time / on another stream and hand the result to forward(value=...)."""
        N, Len_in, _ = input_flatten.shape
        value = dense.linear(input_flatten, self.value_proj.weight, self.value_proj.bias, row_mask=input_padding_mask)
        return value.view(N, Len_in, self.n_heads, self.d_model // self.n_heads)

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes,
                input_level_start_index, input_padding_mask=None, spatial_shapes_host=None, value=None):
        """Arguments as ms_deform_attn.py:82-94.  ``spatial_shapes_host`` (optional, a python list
        of (H, W)) lets callers that already know the level shapes skip the device->host sync the
        reference's ``assert (shapes[:,0]*shapes[:,1]).sum() == Len_in`` costs on every call
        (ms_deform_attn.py:96); without it the assert is evaluated exactly as in the reference."""
        N, Len_q, _ = query.shape
        N, Len_in, _ = input_flatten.shape
        if spatial_shapes_host is not None:
            assert sum(int(h) * int(w) for h, w in spatial_shapes_host) == Len_in
        else:
            assert (input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum() == Len_in

        if value is None:
            value = self.project_value(input_flatten, input_padding_mask)
        if (_FUSED_PROLOGUE and not reference_points.requires_grad and value.is_cuda
                and input_flatten.dtype == torch.float32 and self.d_model // self.n_heads == 32
                and reference_points.shape[-1] in ((2, 4) if _FUSED_PROLOGUE_REF4 else (2,))
                and self.n_levels == 4 and self.n_points == 4
                and value.numel() < 2 ** 32):
            # one GEMM for offsets | logits, then the fused-prologue kernels (encoder self-attention: 2-d reference
            # points; decoder cross-attention from the second layer on: detached 4-d anchors)
            w = torch.cat((self.sampling_offsets.weight, self.attention_weights.weight), 0)
            b = torch.cat((self.sampling_offsets.bias, self.attention_weights.bias), 0)
            proj = dense.linear(query, w, b)
            output = MSDeformAttnProjFunction.apply(value.contiguous(), input_spatial_shapes, input_level_start_index,
                                                    reference_points.contiguous(), proj)
            return dense.linear(output, self.output_proj.weight, self.output_proj.bias)
        sampling_offsets = dense.linear(query, self.sampling_offsets.weight, self.sampling_offsets.bias).view(
            N, Len_q, self.n_heads, self.n_levels, self.n_points, 2)
        attention_weights = dense.linear(query, self.attention_weights.weight, self.attention_weights.bias).view(
            N, Len_q, self.n_heads, self.n_levels * self.n_points)
        attention_weights = F.softmax(attention_weights, -1).view(
            N, Len_q, self.n_heads, self.n_levels, self.n_points)
        if reference_points.shape[-1] == 2:
            offset_normalizer = torch.stack([input_spatial_shapes[..., 1], input_spatial_shapes[..., 0]], -1)
            sampling_locations = reference_points[:, :, None, :, None, :] \
                + sampling_offsets / offset_normalizer[None, None, None, :, None, :]
        elif reference_points.shape[-1] == 4:
            sampling_locations = reference_points[:, :, None, :, None, :2] \
                + sampling_offsets / self.n_points * reference_points[:, :, None, :, None, 2:] * 0.5
        else:
            raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead.".format(
                reference_points.shape[-1]))
        output = MSDeformAttnFunction.apply(value, input_spatial_shapes, input_level_start_index,
                                            sampling_locations, attention_weights, self.im2col_step)
        return dense.linear(output, self.output_proj.weight, self.output_proj.bias)
