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
_lib.rlipv2_dense_linear_tf32_rowmask.argtypes = [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p]
_lib.rlipv2_dense_linear_tf32_rowmask.restype = _i
_lib.rlipv2_dense_linear_tf32_supported.argtypes = [_i, _i, _i]
_lib.rlipv2_dense_linear_tf32_supported.restype = _i
_lib.rlipv2_dense_wgrad_tf32.argtypes = [_p, _p, _p, _i, _i, _i, _i, _p]
_lib.rlipv2_dense_wgrad_tf32.restype = _i
_lib.rlipv2_dense_dgrad_tf32.argtypes = [_p, _p, _p, _p, _p, _i, _i, _i, _p]
_lib.rlipv2_dense_dgrad_tf32.restype = _i
_lib.rlipv2_dense_linear_splitk_tf32.argtypes = [_p, _p, _p, _p, _i, _i, _i, _i, _p]
_lib.rlipv2_dense_linear_splitk_tf32.restype = _i
_lib.rlipv2_dense_error_string.argtypes = [_i]
_lib.rlipv2_dense_error_string.restype = ctypes.c_char_p
_lib.rlipv2_dense_launch_count.restype = ctypes.c_ulonglong
_lib.rlipv2_dense_set_small_mode.argtypes = [_i]
_lib.rlipv2_dense_set_small_mode.restype = None
_lib.rlipv2_dense_get_small_mode.restype = _i
_lib.rlipv2_dense_set_persistent_min_tiles.argtypes = [_i]
_lib.rlipv2_dense_set_persistent_min_tiles.restype = None
_lib.rlipv2_dense_get_persistent_min_tiles.restype = _i
# persistent kernel for linears of more than 2 x 148 tiles (the encoder's 44k-row projections and FFN-up): measured r02n, two
# repetitions on one box: 26.13 / 26.86 vs 26.59 / 27.28 ms/step without.  RLIPV2_DENSE_PERSISTENT=0 turns it off.
_lib.rlipv2_dense_set_persistent_min_tiles(int(os.environ.get("RLIPV2_DENSE_PERSISTENT", "296")))
_lib.rlipv2_dense_set_persistent_dgrad.argtypes = [_i]
_lib.rlipv2_dense_set_persistent_dgrad.restype = None
_lib.rlipv2_dense_get_persistent_dgrad.restype = _i
_lib.rlipv2_dense_set_persistent_dgrad(int(os.environ.get("RLIPV2_DENSE_PERSISTENT_DGRAD", "0")))
if os.environ.get("RLIPV2_DENSE_SMALL_MODE"):                      # A/B switch for measurements
    _lib.rlipv2_dense_set_small_mode(int(os.environ["RLIPV2_DENSE_SMALL_MODE"]))

ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
EXPORTS = ("rlipv2_dense_linear_tf32", "rlipv2_dense_linear_tf32_rowmask", "rlipv2_dense_linear_tf32_supported", "rlipv2_dense_wgrad_tf32",
           "rlipv2_dense_dgrad_tf32", "rlipv2_dense_error_string", "rlipv2_dense_launch_count",
           "rlipv2_dense_set_small_mode", "rlipv2_dense_get_small_mode", "rlipv2_dense_linear_splitk_tf32",
           "rlipv2_dense_set_persistent_min_tiles", "rlipv2_dense_get_persistent_min_tiles",
           "rlipv2_dense_set_persistent_dgrad", "rlipv2_dense_get_persistent_dgrad")


def set_persistent_dgrad(on):
    _lib.rlipv2_dense_set_persistent_dgrad(1 if on else 0)


def set_persistent_min_tiles(tiles):
    """linears with more than `tiles` 128 x 128 output tiles run the persistent kernel; 0 = never (include/rlipv2_dense.h)"""
    _lib.rlipv2_dense_set_persistent_min_tiles(int(tiles))


def persistent_min_tiles():
    return int(_lib.rlipv2_dense_get_persistent_min_tiles())


def set_small_mode(mode):
    """tile / pipeline choice of the forward linear for grids of at most one CTA per SM (include/rlipv2_dense.h)"""
    _lib.rlipv2_dense_set_small_mode(int(mode))


def small_mode():
    return int(_lib.rlipv2_dense_get_small_mode())


def library_path():
    return _path


def launch_count():
    return int(_lib.rlipv2_dense_launch_count())


def supported(M, N, K):
    return bool(_lib.rlipv2_dense_linear_tf32_supported(M, N, K))


def linear_tf32(x2d, weight, bias, act=ACT_NONE, rowmask=None):
    """x2d [M,K], weight [N,K], bias [N]|None - contiguous fp32 CUDA tensors -> y [M,N]; rows r with rowmask[r]
    (bool [M], contiguous) come out as zeros"""
    M, K = x2d.shape
    N = weight.shape[0]
    y = torch.empty((M, N), dtype=torch.float32, device=x2d.device)
    with torch.cuda.device(x2d.device):
        rc = _lib.rlipv2_dense_linear_tf32_rowmask(
            x2d.data_ptr(), weight.data_ptr(), bias.data_ptr() if bias is not None else None,
            rowmask.data_ptr() if rowmask is not None else None, y.data_ptr(), M, N, K, act,
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise RuntimeError(f"rlipv2_dense_linear_tf32: {_lib.rlipv2_dense_error_string(rc).decode()} (code {rc})")
    return y


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def splitk_splits(M, N, K):
    """K slices per output tile for the split-K forward linear: fill one wave of 148 SMs with (tiles x slices) CTAs, at
    least 4 k-blocks of 32 per slice; 1 = not worth splitting (use the plain kernel)"""
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    kb = K // 32
    if K % 32 or N % 4 or tiles > 74 or K < 768:          # (short K: the zero-fill + reductions cost more than they buy)
        return 1
    return max(1, min(148 // tiles, kb // 4))


def linear_splitk_tf32(x2d, weight, bias, splits):
    """y [M,N] = x2d [M,K] @ weight[N,K]^T + bias, K split over `splits` CTAs per output tile (fp32 reductions into y)"""
    M, K = x2d.shape
    N = weight.shape[0]
    y = torch.empty((M, N), dtype=torch.float32, device=x2d.device)
    with torch.cuda.device(x2d.device):
        rc = _lib.rlipv2_dense_linear_splitk_tf32(x2d.data_ptr(), weight.data_ptr(),
                                                  bias.data_ptr() if bias is not None else None, y.data_ptr(), M, N, K,
                                                  int(splits), _stream())
    if rc != 0:
        raise RuntimeError(f"rlipv2_dense_linear_splitk_tf32: {_lib.rlipv2_dense_error_string(rc).decode()} (code {rc})")
    return y


def grads_supported(T, N, K):
    """shapes the tcgen05 backward GEMMs take (row strides must be 16-byte multiples for TMA)"""
    return T > 0 and N > 0 and K > 0 and N % 4 == 0 and K % 4 == 0


def wgrad_splits(T, N, K):
    """CTAs along the token axis per output tile: fill ~2 waves of 148 SMs, at least 8 k-blocks of 32 rows each"""
    bn = 256 if K % 256 == 0 else 128
    tiles = ((N + 127) // 128) * ((K + bn - 1) // bn)
    kb = (T + 31) // 32
    # measured on the encoder's 44k-token shapes (profiles/dense_bwd_microbench_r01.jsonl): one wave of CTAs, never
    # more (a 2048x256 weight at 10 splits = 160 CTAs runs 45 % slower than at 8 = 128 CTAs on 148 SMs)
    return max(1, min(148 // tiles if tiles <= 148 else 1, kb // 8 if kb >= 8 else 1))


def wgrad_tf32(g2d, x2d, splits=None, acc=None):
    """dw [N,K] = g2d[T,N]^T @ x2d[T,K] (contiguous fp32 CUDA), split over the T axis and reduced with fp32 reductions.
    `acc` (contiguous fp32 [N,K], e.g. the weight's view of the flat gradient buffer): the product is ADDED into it (the
    kernel reduces with red.global.add anyway) instead of into a fresh zero-filled tensor."""
    T, N = g2d.shape
    K = x2d.shape[1]
    if acc is not None:
        assert acc.shape == (N, K) and acc.is_contiguous() and acc.dtype == torch.float32 and acc.is_cuda
        dw = acc
    else:
        dw = torch.zeros((N, K), dtype=torch.float32, device=g2d.device)
    with torch.cuda.device(g2d.device):
        rc = _lib.rlipv2_dense_wgrad_tf32(g2d.data_ptr(), x2d.data_ptr(), dw.data_ptr(), T, N, K,
                                          splits if splits is not None else wgrad_splits(T, N, K), _stream())
    if rc != 0:
        raise RuntimeError(f"rlipv2_dense_wgrad_tf32: {_lib.rlipv2_dense_error_string(rc).decode()} (code {rc})")
    return dw


def dgrad_tf32(g2d, weight, relu_out=None):
    """dx [T,K] = g2d[T,N] @ weight[N,K]; with `relu_out` [T,K]: dx masked by (relu_out > 0) and its column sums
    -> (dx, colsum | None)"""
    T, N = g2d.shape
    K = weight.shape[1]
    dx = torch.empty((T, K), dtype=torch.float32, device=g2d.device)
    colsum = torch.empty(((T + 127) // 128, K), dtype=torch.float32, device=g2d.device) if relu_out is not None else None
    with torch.cuda.device(g2d.device):
        rc = _lib.rlipv2_dense_dgrad_tf32(g2d.data_ptr(), weight.data_ptr(), dx.data_ptr(),
                                          relu_out.data_ptr() if relu_out is not None else None,
                                          colsum.data_ptr() if colsum is not None else None, T, N, K, _stream())
    if rc != 0:
        raise RuntimeError(f"rlipv2_dense_dgrad_tf32: {_lib.rlipv2_dense_error_string(rc).decode()} (code {rc})")
    if colsum is not None:                       # add the row-tile partials up
        if K % 32 == 0:
            from . import fused_abi
            colsum = fused_abi.relu_bwd_colsum(colsum, None)[1]
        else:
            colsum = colsum.sum(0)
    return dx, colsum
