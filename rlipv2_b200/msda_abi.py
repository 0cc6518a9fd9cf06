"""ctypes binding of the C ABI declared in include/rlipv2_msda.h.

PyTorch is used only for device memory and the current CUDA stream; every tensor crosses the
boundary as a raw device pointer.  There is no fallback: if the CUDA library has not been built
(``python -m rlipv2_b200.build``) importing this module raises.
"""
import ctypes
import os

import torch

from .build import lib_path

_LIB_NAME = "librlipv2_msda.so"
ABI_VERSION = 2

_path = lib_path(_LIB_NAME)
if not os.path.exists(_path):
    raise ImportError(
        f"{_path} is missing: the sm_100a CUDA library has not been built "
        "(run `python -m rlipv2_b200.build` or `__graft_entry__.build()`); "
        "rlipv2_b200 has no CPU or PyTorch fallback for MSDeformAttn.")
_lib = ctypes.CDLL(_path)

_i, _p = ctypes.c_int, ctypes.c_void_p
_DIMS = [_i] * 7
_lib.rlipv2_msda_forward_f32.argtypes = [_p] * 5 + _DIMS + [_p, _p]
_lib.rlipv2_msda_forward_f64.argtypes = [_p] * 5 + _DIMS + [_p, _p]
_lib.rlipv2_msda_backward_f32.argtypes = [_p] * 6 + _DIMS + [_p] * 4
_lib.rlipv2_msda_backward_f64.argtypes = [_p] * 6 + _DIMS + [_p] * 4
_lib.rlipv2_msda_proj_forward_f32.argtypes = [_p] * 5 + _DIMS + [_p, _p]
_lib.rlipv2_msda_proj_backward_f32.argtypes = [_p] * 6 + _DIMS + [_p] * 3
_lib.rlipv2_msda_proj_ref4_forward_f32.argtypes = [_p] * 5 + _DIMS + [_p, _p]
_lib.rlipv2_msda_proj_ref4_backward_f32.argtypes = [_p] * 6 + _DIMS + [_p] * 3
_lib.rlipv2_msda_forward_tma_f32.argtypes = [_p] * 5 + _DIMS + [_i, _p, _p]
_lib.rlipv2_msda_forward_tma_f32.restype = _i
for _f in ("forward_f32", "forward_f64", "backward_f32", "backward_f64", "proj_forward_f32", "proj_backward_f32",
           "proj_ref4_forward_f32", "proj_ref4_backward_f32"):
    getattr(_lib, "rlipv2_msda_" + _f).restype = _i
_lib.rlipv2_msda_error_string.argtypes = [_i]
_lib.rlipv2_msda_error_string.restype = ctypes.c_char_p
_lib.rlipv2_msda_abi_version.restype = _i
_lib.rlipv2_msda_launch_count.restype = ctypes.c_ulonglong
_lib.rlipv2_msda_set_backward_mode.argtypes = [_i]
_lib.rlipv2_msda_set_backward_mode.restype = _i
_lib.rlipv2_msda_get_backward_mode.restype = _i

if _lib.rlipv2_msda_abi_version() != ABI_VERSION:
    raise ImportError(f"{_path}: ABI version {_lib.rlipv2_msda_abi_version()} != {ABI_VERSION}; rebuild")

EXPORTS = ("rlipv2_msda_forward_f32", "rlipv2_msda_forward_f64", "rlipv2_msda_backward_f32",
           "rlipv2_msda_backward_f64", "rlipv2_msda_proj_forward_f32", "rlipv2_msda_proj_backward_f32",
           "rlipv2_msda_proj_ref4_forward_f32", "rlipv2_msda_proj_ref4_backward_f32",
           "rlipv2_msda_forward_tma_f32", "rlipv2_msda_error_string", "rlipv2_msda_abi_version", "rlipv2_msda_launch_count",
           "rlipv2_msda_set_backward_mode", "rlipv2_msda_get_backward_mode")


def library_path():
    return _path


def set_backward_mode(mode):
    """0 = one reduction per valid corner, 1 = same-cell corners of a pair merged before issue, 2 (the library's default) =
    chosen by call shape, 3 / 4 = measurement variants (include/rlipv2_msda.h)"""
    _check(_lib.rlipv2_msda_set_backward_mode(int(mode)), "rlipv2_msda_set_backward_mode")


def get_backward_mode():
    return int(_lib.rlipv2_msda_get_backward_mode())



def launch_count():
    return int(_lib.rlipv2_msda_launch_count())


def _check(code, what):
    if code != 0:
        raise RuntimeError(f"{what}: {_lib.rlipv2_msda_error_string(code).decode()} (code {code})")


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, out):
    """Raw call: all tensors CUDA, contiguous, same floating dtype; `out` preallocated."""
    N, S, M, D = value.shape
    Lq, L, P = sampling_loc.shape[1], sampling_loc.shape[3], sampling_loc.shape[4]
    fn = _lib.rlipv2_msda_forward_f32 if value.dtype == torch.float32 else _lib.rlipv2_msda_forward_f64
    with torch.cuda.device(value.device):
        code = fn(value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                  sampling_loc.data_ptr(), attn_weight.data_ptr(), N, S, M, D, L, Lq, P,
                  out.data_ptr(), _stream())
    _check(code, "ms_deform_attn_forward")


def forward_tma(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, coarse_start):
    """Experimental forward with the two coarsest levels staged in shared memory by TMA (include/rlipv2_msda.h);
    coarse_start = level_start_index[-2] as a host int.  fp32, D = 32, L = 4, P = 4 only."""
    N, S, M, D = value.shape
    Lq, L, P = sampling_loc.shape[1], sampling_loc.shape[3], sampling_loc.shape[4]
    out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
    with torch.cuda.device(value.device):
        code = _lib.rlipv2_msda_forward_tma_f32(value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                                                sampling_loc.data_ptr(), attn_weight.data_ptr(), N, S, M, D, L, Lq, P,
                                                int(coarse_start), out.data_ptr(), _stream())
    _check(code, "rlipv2_msda_forward_tma_f32")
    return out


def backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
             grad_value, grad_sampling_loc, grad_attn_weight):
    N, S, M, D = value.shape
    Lq, L, P = sampling_loc.shape[1], sampling_loc.shape[3], sampling_loc.shape[4]
    fn = _lib.rlipv2_msda_backward_f32 if value.dtype == torch.float32 else _lib.rlipv2_msda_backward_f64
    with torch.cuda.device(value.device):
        code = fn(value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                  sampling_loc.data_ptr(), attn_weight.data_ptr(), grad_output.data_ptr(),
                  N, S, M, D, L, Lq, P, grad_value.data_ptr(), grad_sampling_loc.data_ptr(),
                  grad_attn_weight.data_ptr(), _stream())
    _check(code, "ms_deform_attn_backward")


def proj_supported(value, reference_points, n_levels, n_points):
    """shapes the fused-prologue entry points accept (fp32, D=32, L=4, P=4, 2-d or 4-d reference points)"""
    return (value.is_cuda and value.dtype == torch.float32 and value.shape[-1] == 32 and n_levels == 4
            and n_points == 4 and reference_points.shape[-1] in (2, 4) and value.numel() < 2 ** 32)


def proj_forward(value, spatial_shapes, level_start_index, reference_points, proj, out):
    """value [N,S,M,32], reference_points [N,Lq,4,2|4], proj [N,Lq,M*48] (offsets | logits) -> out [N,Lq,M*32]"""
    N, S, M, D = value.shape
    Lq = proj.shape[1]
    fn = _lib.rlipv2_msda_proj_ref4_forward_f32 if reference_points.shape[-1] == 4 else _lib.rlipv2_msda_proj_forward_f32
    with torch.cuda.device(value.device):
        code = fn(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), reference_points.data_ptr(),
            proj.data_ptr(), N, S, M, D, 4, Lq, 4, out.data_ptr(), _stream())
    _check(code, "msda_proj_forward")


def proj_backward(value, spatial_shapes, level_start_index, reference_points, proj, grad_output, grad_value, grad_proj):
    N, S, M, D = value.shape
    Lq = proj.shape[1]
    fn = _lib.rlipv2_msda_proj_ref4_backward_f32 if reference_points.shape[-1] == 4 else _lib.rlipv2_msda_proj_backward_f32
    with torch.cuda.device(value.device):
        code = fn(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), reference_points.data_ptr(),
            proj.data_ptr(), grad_output.data_ptr(), N, S, M, D, 4, Lq, 4, grad_value.data_ptr(),
            grad_proj.data_ptr(), _stream())
    _check(code, "msda_proj_backward")


if os.environ.get("RLIPV2_MSDA_BWD_MERGE"):                            # A/B switch for measurements (0 / 1 / 2)
    set_backward_mode(int(os.environ["RLIPV2_MSDA_BWD_MERGE"]))
