"""Drop-in for the reference's native extension module ``MultiScaleDeformableAttention``.

The reference imports it as ``import MultiScaleDeformableAttention as MSDA``
(/root/reference/models/ops/functions/ms_deform_attn_func.py:22 and the byte-identical copy under
models/dab_deformable/ops/) and calls the two functions pybind11 exports in
models/ops/src/vision.cpp:13-16.  Put ``rlipv2_b200/dropin`` on ``PYTHONPATH`` ahead of any other
build of that module and the reference's ``MSDeformAttnFunction`` / ``MSDeformAttn`` run on the
sm_100a kernels unchanged.

Signatures, preconditions and error behaviour follow ms_deform_attn.h:20-61 and
ms_deform_attn_cuda.cu:20-153:
  * every tensor must be contiguous and CUDA, else RuntimeError (AT_ASSERTM, cu:28-38,93-105);
    a CPU ``value`` raises "Not implemented on the CPU" (ms_deform_attn.h:35,60);
  * ``batch % min(batch, im2col_step) == 0`` (cu:52,119);
  * float32 and float64 (AT_DISPATCH_FLOATING_TYPES, cu:64,134);
  * outputs are freshly allocated: forward [N, Lq, M*D]; backward [grad_value,
    grad_sampling_loc, grad_attn_weight] shaped like their inputs;
  * launched on the current CUDA stream, no host synchronisation.
"""
import torch

from rlipv2_b200 import msda_abi as _abi

__all__ = ["ms_deform_attn_forward", "ms_deform_attn_backward"]


def _validate(named, im2col_step):
    value = named[0][1]
    if not value.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    for name, t in named:
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor")
    if value.dtype not in (torch.float32, torch.float64):
        raise RuntimeError(f'"ms_deform_attn" not implemented for \'{value.dtype}\'')
    for name, t in named:
        want = torch.int64 if name in ("spatial_shapes", "level_start_index") else value.dtype
        if t.dtype != want:     # Tensor::data<scalar_t>() throws on a dtype mismatch (cu:66-71)
            raise RuntimeError(f"expected scalar type {want} but found {t.dtype} for {name}")
    batch = value.size(0)
    step = min(batch, int(im2col_step))
    if batch > 0 and (step <= 0 or batch % step != 0):
        raise RuntimeError(f"batch({batch}) must divide im2col_step({step})")


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                           im2col_step):
    _validate([("value", value), ("spatial_shapes", spatial_shapes),
               ("level_start_index", level_start_index), ("sampling_loc", sampling_loc),
               ("attn_weight", attn_weight)], im2col_step)
    N, _, M, D = value.shape
    Lq = sampling_loc.size(1)
    out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
    _abi.forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, out)
    return out


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                            grad_output, im2col_step):
    _validate([("value", value), ("spatial_shapes", spatial_shapes),
               ("level_start_index", level_start_index), ("sampling_loc", sampling_loc),
               ("attn_weight", attn_weight), ("grad_output", grad_output)], im2col_step)
    grad_value = torch.empty_like(value)          # zero-filled inside the library call
    grad_sampling_loc = torch.empty_like(sampling_loc)
    grad_attn_weight = torch.empty_like(attn_weight)
    _abi.backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                  grad_output, grad_value, grad_sampling_loc, grad_attn_weight)
    return [grad_value, grad_sampling_loc, grad_attn_weight]
