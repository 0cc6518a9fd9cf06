"""ctypes binding of include/rlipv2_lsap.h (the matcher's assignment problems on the device).  No fallback: a missing
library raises."""
import ctypes
import os

import torch

from .build import lib_path

_path = lib_path("librlipv2_lsap.so")
if not os.path.exists(_path):
    raise ImportError(f"{_path} is missing: run `python -m rlipv2_b200.build` (no CPU / PyTorch fallback is provided)")
_lib = ctypes.CDLL(_path)
_i, _p = ctypes.c_int, ctypes.c_void_p
_lib.rlipv2_lsap_f32.argtypes = [_p, _i, _i, _i, _i, _p, _p, _i, _p, _p, _p, _p, _p]
_lib.rlipv2_lsap_f32.restype = _i
_lib.rlipv2_lsap_error_string.argtypes = [_i]
_lib.rlipv2_lsap_error_string.restype = ctypes.c_char_p
_lib.rlipv2_lsap_launch_count.restype = ctypes.c_ulonglong

EXPORTS = ("rlipv2_lsap_f32", "rlipv2_lsap_error_string", "rlipv2_lsap_launch_count")


def library_path():
    return _path


def launch_count():
    return int(_lib.rlipv2_lsap_launch_count())


class Plan:
    """Device-side description of a batch's assignment problems, built once per target layout (the host knows the
    number of ground-truth triplets per image; nothing here syncs with the device after construction).

    sizes: triplets per image; nq: queries per image; n_levels: decoder levels solved together.
    Output layout = criterion.StackedMatches: pairs ordered (level, image, match), min(nq, sizes[b]) per problem."""

    def __init__(self, sizes, nq, n_levels, device):
        self.sizes, self.nq, self.n_levels = [int(s) for s in sizes], int(nq), int(n_levels)
        starts, o = [], 0
        for s in self.sizes:
            starts.append(o)
            o += s
        self.T = o
        self.ks = [min(self.nq, s) for s in self.sizes]
        offs, o = [], 0
        for _ in range(self.n_levels):
            for k in self.ks:
                offs.append(o)
                o += k
        self.K = o
        self.max_count = max(self.sizes) if self.sizes else 0
        self.tgt_start = torch.tensor(starts, dtype=torch.int32, device=device)
        self.tgt_count = torch.tensor(self.sizes, dtype=torch.int32, device=device)
        self.out_offset = torch.tensor(offs, dtype=torch.int64, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)


def solve(cost, plan, out_query=None, out_target=None):
    """cost [n_levels, bs, nq, T] fp32 CUDA (contiguous) -> (query_idx [K], target_idx [K]) int64 CUDA tensors in
    `plan`'s stacked layout; equals scipy.optimize.linear_sum_assignment per (level, image) bit for bit.
    Asynchronous on the current stream; `check(plan)` (a device->host read) reports problems scipy would have refused."""
    if not (cost.is_cuda and cost.dtype == torch.float32 and cost.is_contiguous()):
        raise RuntimeError("lsap.solve: cost must be a contiguous fp32 CUDA tensor")
    if tuple(cost.shape) != (plan.n_levels, len(plan.sizes), plan.nq, plan.T):
        raise RuntimeError(f"lsap.solve: cost shape {tuple(cost.shape)} does not match the plan "
                           f"{(plan.n_levels, len(plan.sizes), plan.nq, plan.T)}")
    if out_query is None:
        out_query = torch.empty(plan.K, dtype=torch.int64, device=cost.device)
        out_target = torch.empty(plan.K, dtype=torch.int64, device=cost.device)
    elif not (out_query.is_cuda and out_target.is_cuda and out_query.dtype == out_target.dtype == torch.int64
              and out_query.numel() == out_target.numel() == plan.K and out_query.is_contiguous()
              and out_target.is_contiguous()):
        raise RuntimeError("lsap.solve: output buffers must be contiguous int64 CUDA tensors of plan.K elements")
    rc = _lib.rlipv2_lsap_f32(cost.data_ptr(), plan.n_levels, len(plan.sizes), plan.nq, plan.T, plan.tgt_start.data_ptr(),
                              plan.tgt_count.data_ptr(), plan.max_count, plan.out_offset.data_ptr(), out_query.data_ptr(),
                              out_target.data_ptr(), plan.err.data_ptr(),
                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise RuntimeError(f"rlipv2_lsap_f32: {_lib.rlipv2_lsap_error_string(rc).decode()} (code {rc})")
    return out_query, out_target


def check(plan):
    e = int(plan.err.item())
    if e:
        raise ValueError(f"assignment problem {e - 1} (level * batch + image) is infeasible or holds non-finite costs")
