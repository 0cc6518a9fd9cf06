"""ctypes binding of include/rlipv2_dense.h (tcgen05 TF32 linear).  No fallback: a missing library raises."""
import ctypes
import os

import torch

from .build import lib_path

_path = lib_path("librlipv2_dense.so")
if not os.path.exists(_path):
    raise ImportError(f"{_path} is missing: run `python -m rlipv2_b200.build` (no CPU / PyTorch fallback is provided)")
_lib = ctypes.CDLL(_path)
_i, _p = ctypes.c_int, ctypes.c_void_p
_lib.rlipv2_dense_linear_tf32.argtypes = [_p, _p, _p, _p, _i, _i, _i, _i, _p]
_lib.rlipv2_dense_linear_tf32.restype = _i
_lib.rlipv2_dense_linear_tf32_supported.argtypes = [_i, _i, _i]
_lib.rlipv2_dense_linear_tf32_supported.restype = _i
_lib.rlipv2_dense_error_string.argtypes = [_i]
_lib.rlipv2_dense_error_string.restype = ctypes.c_char_p
_lib.rlipv2_dense_launch_count.restype = ctypes.c_ulonglong

ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
EXPORTS = ("rlipv2_dense_linear_tf32", "rlipv2_dense_linear_tf32_supported", "rlipv2_dense_error_string",
           "rlipv2_dense_launch_count")


def library_path():
    return _path


def launch_count():
    return int(_lib.rlipv2_dense_launch_count())


def supported(M, N, K):
    return bool(_lib.rlipv2_dense_linear_tf32_supported(M, N, K))


def linear_tf32(x2d, weight, bias, act=ACT_NONE):
    """x2d [M,K], weight [N,K], bias [N]|None - contiguous fp32 CUDA tensors -> y [M,N]"""
    M, K = x2d.shape
    N = weight.shape[0]
    y = torch.empty((M, N), dtype=torch.float32, device=x2d.device)
    with torch.cuda.device(x2d.device):
        rc = _lib.rlipv2_dense_linear_tf32(x2d.data_ptr(), weight.data_ptr(), bias.data_ptr() if bias is not None else None,
                                           y.data_ptr(), M, N, K, act,
                                           ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise RuntimeError(f"rlipv2_dense_linear_tf32: {_lib.rlipv2_dense_error_string(rc).decode()} (code {rc})")
    return y
