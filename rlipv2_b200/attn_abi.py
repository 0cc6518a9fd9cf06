"""ctypes binding of include/rlipv2_attn.h (fused tcgen05 attention cores).  No fallback: a missing library raises."""
import ctypes
import os

import torch

from .build import lib_path

_path = lib_path("librlipv2_attn.so")
if not os.path.exists(_path):
    raise ImportError(f"{_path} is missing: run `python -m rlipv2_b200.build` (no CPU / PyTorch fallback is provided)")
_lib = ctypes.CDLL(_path)
_i, _p, _ll, _f, _d, _u = ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_float, ctypes.c_double, ctypes.c_uint
_lib.rlipv2_attn_supported.argtypes = [_i] * 5
_lib.rlipv2_attn_supported.restype = _i
_lib.rlipv2_attn_key_pitch.argtypes = [_i]
_lib.rlipv2_attn_key_pitch.restype = _i
_lib.rlipv2_attn_forward_tf32.argtypes = ([_p, _ll, _ll] * 3 + [_p, _p, _ll, _ll, _p] + [_i] * 5 + [_f, _d, _p, _u, _p, _p])
_lib.rlipv2_attn_forward_tf32.restype = _i
_lib.rlipv2_attn_backward_tf32.argtypes = ([_p, _ll, _ll] * 3 + [_p, _p, _p, _ll, _ll, _p] + [_p, _ll, _ll] * 3 + [_p, _p]
                                           + [_i] * 5 + [_f, _d, _p, _u, _i, _p])
_lib.rlipv2_attn_backward_tf32.restype = _i
_lib.rlipv2_attn_error_string.argtypes = [_i]
_lib.rlipv2_attn_error_string.restype = ctypes.c_char_p
_lib.rlipv2_attn_launch_count.restype = ctypes.c_ulonglong

EXPORTS = ("rlipv2_attn_supported", "rlipv2_attn_key_pitch", "rlipv2_attn_forward_tf32", "rlipv2_attn_backward_tf32",
           "rlipv2_attn_error_string", "rlipv2_attn_launch_count")


def library_path():
    return _path


def launch_count():
    return int(_lib.rlipv2_attn_launch_count())


def supported(B, H, Tq, Nk, D):
    return bool(_lib.rlipv2_attn_supported(B, H, Tq, Nk, D))


def key_pitch(Nk):
    return int(_lib.rlipv2_attn_key_pitch(Nk))


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what}: {_lib.rlipv2_attn_error_string(rc).decode()} (code {rc})")


def _desc(t):
    """[B, T, H*D] fp32 CUDA tensor with unit stride along the last dim -> (pointer, row stride, batch stride)"""
    assert t.is_cuda and t.dtype == torch.float32 and t.dim() == 3 and t.stride(2) == 1
    return t.data_ptr(), t.stride(1), t.stride(0)


def usable(t):
    """can `t` [B, T, C] be handed to the kernels in place (unit inner stride, 16-byte aligned rows)?"""
    return (t.dim() == 3 and t.stride(2) == 1 and t.stride(1) % 4 == 0 and t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0)


def forward(q, k, v, heads, key_bias, scale, dropout_p, seed, salt):
    """q [B, Tq, H*D], k / v [B, Nk, H*D] (views allowed, see `usable`); key_bias [B, Nk] contiguous or None
    -> (out [B, Tq, H*D], stats [B*H, Tq, 2], seed_used int64 [1])"""
    B, Tq, C = q.shape
    Nk = k.shape[1]
    D = C // heads
    out = torch.empty((B, Tq, C), dtype=torch.float32, device=q.device)
    stats = torch.empty((B * heads, Tq, 2), dtype=torch.float32, device=q.device)
    seed_used = torch.zeros(1, dtype=torch.int64, device=q.device)
    with torch.cuda.device(q.device):
        rc = _lib.rlipv2_attn_forward_tf32(*_desc(q), *_desc(k), *_desc(v),
                                           key_bias.data_ptr() if key_bias is not None else None, out.data_ptr(),
                                           out.stride(1), out.stride(0), stats.data_ptr(), B, heads, Tq, Nk, D, float(scale),
                                           float(dropout_p), seed.data_ptr() if (seed is not None and dropout_p > 0) else None,
                                           int(salt) & 0xFFFFFFFF, seed_used.data_ptr(), _stream())
    _check(rc, "rlipv2_attn_forward_tf32")
    return out, stats, seed_used


def backward(q, k, v, heads, key_bias, out, dout, stats, scale, dropout_p, seed_used, salt, dq=None, dk=None, dv=None):
    """-> (dq, dk, dv) contiguous [B, T, H*D]; given dq / dk / dv (contiguous, pre-initialised) are ACCUMULATED into"""
    B, Tq, C = q.shape
    Nk = k.shape[1]
    D = C // heads
    accumulate = dq is not None
    if accumulate:
        assert dk is not None and dv is not None
    else:
        dq = torch.empty((B, Tq, C), dtype=torch.float32, device=q.device)
        dk = torch.empty((B, Nk, C), dtype=torch.float32, device=q.device)
        dv = torch.empty((B, Nk, C), dtype=torch.float32, device=q.device)
    pitch = key_pitch(Nk)
    ws = torch.empty((2, B * heads, Tq, pitch), dtype=torch.float32, device=q.device)
    assert out.stride() == dout.stride()
    with torch.cuda.device(q.device):
        rc = _lib.rlipv2_attn_backward_tf32(*_desc(q), *_desc(k), *_desc(v),
                                            key_bias.data_ptr() if key_bias is not None else None, out.data_ptr(),
                                            dout.data_ptr(), out.stride(1), out.stride(0), stats.data_ptr(), *_desc(dq),
                                            *_desc(dk), *_desc(dv), ws[0].data_ptr(), ws[1].data_ptr(), B, heads, Tq, Nk, D,
                                            float(scale), float(dropout_p), seed_used.data_ptr(), int(salt) & 0xFFFFFFFF,
                                            1 if accumulate else 0, _stream())
    _check(rc, "rlipv2_attn_backward_tf32")
    return dq, dk, dv
