"""TEST INFRASTRUCTURE - NOT PRODUCT CODE.

numpy/ctypes front end of the C oracle ``oracle/msda_oracle.c`` (a restatement of
/root/reference/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:34-159,238-403).

Pinned against the reference's own ``ms_deform_attn_core_pytorch``
(models/ops/functions/ms_deform_attn_func.py:47-65) by tests/test_oracle_msda.py via
the committed fixtures tests/golden/msda_*.npz (generator: oracle/gen_golden_msda.py).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libmsda_oracle.so")
_lib = None


def build(force=False):
    """Compile the C oracle with gcc (seconds)."""
    src = [os.path.join(_HERE, f) for f in ("msda_oracle.c", "msda_oracle_impl.h", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _prep(value, shapes, lsi, loc, attn):
    dt = np.float64 if value.dtype == np.float64 else np.float32
    value = np.ascontiguousarray(value, dtype=dt)
    loc = np.ascontiguousarray(loc, dtype=dt)
    attn = np.ascontiguousarray(attn, dtype=dt)
    shapes = np.ascontiguousarray(shapes, dtype=np.int64)
    lsi = np.ascontiguousarray(lsi, dtype=np.int64)
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    assert attn.shape == (N, Lq, M, L, P) and shapes.shape == (L, 2) and lsi.shape == (L,)
    return dt, value, shapes, lsi, loc, attn, (N, S, M, D, L, Lq, P)


def forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, threads=1):
    """-> out [N, Lq, M*D] (same dtype as value: float32 or float64)."""
    dt, value, shapes, lsi, loc, attn, (N, S, M, D, L, Lq, P) = _prep(
        value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    out = np.zeros((N, Lq, M * D), dtype=dt)
    lib = _load()
    dims = [ctypes.c_int(x) for x in (N, S, M, D, L, Lq, P)]
    if dt == np.float32 and threads > 1:
        lib.msda_oracle_forward_f32_mt(_ptr(value), _ptr(shapes), _ptr(lsi), _ptr(loc), _ptr(attn),
                                       *dims, _ptr(out), ctypes.c_int(threads))
    else:
        fn = lib.msda_oracle_forward_f32 if dt == np.float32 else lib.msda_oracle_forward_f64
        fn(_ptr(value), _ptr(shapes), _ptr(lsi), _ptr(loc), _ptr(attn), *dims, _ptr(out))
    return out


def backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_out, threads=1):
    """-> (grad_value [N,S,M,D], grad_sampling_loc [N,Lq,M,L,P,2], grad_attn_weight [N,Lq,M,L,P])."""
    dt, value, shapes, lsi, loc, attn, (N, S, M, D, L, Lq, P) = _prep(
        value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    gout = np.ascontiguousarray(grad_out, dtype=dt).reshape(N, Lq, M * D)
    gv = np.zeros_like(value)
    gl = np.zeros_like(loc)
    ga = np.zeros_like(attn)
    lib = _load()
    dims = [ctypes.c_int(x) for x in (N, S, M, D, L, Lq, P)]
    if dt == np.float32 and threads > 1:
        lib.msda_oracle_backward_f32_mt(_ptr(value), _ptr(shapes), _ptr(lsi), _ptr(loc), _ptr(attn),
                                        _ptr(gout), *dims, _ptr(gv), _ptr(gl), _ptr(ga),
                                        ctypes.c_int(threads))
    else:
        fn = lib.msda_oracle_backward_f32 if dt == np.float32 else lib.msda_oracle_backward_f64
        fn(_ptr(value), _ptr(shapes), _ptr(lsi), _ptr(loc), _ptr(attn), _ptr(gout), *dims,
           _ptr(gv), _ptr(gl), _ptr(ga))
    return gv, gl, ga
