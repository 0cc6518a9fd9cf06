"""ctypes binding of include/rlipv2_fused.h (HBM-bound fused kernels).  No fallback: a missing library raises."""
import ctypes
import os

import torch

from .build import lib_path

_path = lib_path("librlipv2_fused.so")
if not os.path.exists(_path):
    raise ImportError(f"{_path} is missing: run `python -m rlipv2_b200.build` (no CPU / PyTorch fallback is provided)")
_lib = ctypes.CDLL(_path)
_i, _p, _f, _ll = ctypes.c_int, ctypes.c_void_p, ctypes.c_float, ctypes.c_longlong
_lib.rlipv2_add_layernorm_fwd_f32.argtypes = [_p, _p, _p, _p, _f, _i, _i, _p, _p, _p, _p, _p]
_lib.rlipv2_layernorm_bwd_f32.argtypes = [_p, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p]
_lib.rlipv2_relu_bwd_colsum_f32.argtypes = [_p, _p, _p, _p, _i, _i, _p]
_d = ctypes.c_double
_lib.rlipv2_adamw_f32.argtypes = [_p, _p, _p, _p, _ll, _d, _d, _d, _d, _d, _p, _p]
_lib.rlipv2_adamw_scaled_f32.argtypes = [_p, _p, _p, _p, _ll, _d, _d, _d, _d, _d, _p, _p, _p]
_lib.rlipv2_adamw_scaled_f32.restype = _i
_lib.rlipv2_adamw_dev_f32.argtypes = [_p, _p, _p, _p, _ll, _d, _p, _d, _d, _d, _d, _p, _p, _p, _p]
_lib.rlipv2_adamw_dev_f32.restype = _i
_lib.rlipv2_gather_chunks_f32.argtypes = [_p, _i, _p, _p]
_lib.rlipv2_rowmask_bwd_colsum_f32.argtypes = [_p, _p, _p, _p, _i, _i, _p]
_ull, _u64p = ctypes.c_ulonglong, ctypes.c_void_p
_lib.rlipv2_wait_host_flag.argtypes = [_p, _p, _ull, _p, _p]
_lib.rlipv2_layernorm_bwd_acc_f32.argtypes = [_p, _p, _p, _p, _p, _i, _i, _p, _p, _p, _i, _p]
_lib.rlipv2_relu_bwd_colsum_acc_f32.argtypes = [_p, _p, _p, _p, _i, _i, _i, _p]
_lib.rlipv2_rowmask_bwd_colsum_acc_f32.argtypes = [_p, _p, _p, _p, _i, _i, _i, _p]
for _n in ("layernorm_bwd_acc_f32", "relu_bwd_colsum_acc_f32", "rowmask_bwd_colsum_acc_f32"):
    getattr(_lib, "rlipv2_" + _n).restype = _i
_lib.rlipv2_stamp_globaltimer.argtypes = [_p, _p]
_lib.rlipv2_stamp_globaltimer.restype = _i
_lib.rlipv2_box_refine_f32.argtypes = [_p, _p, _f, _ll, _p, _p]
_lib.rlipv2_sine_embed_f32.argtypes = [_p, _i, _i, _p, _p]
_lib.rlipv2_box_pair_loss_f32.argtypes = [_p, _p, _i, _p, _p, _p, _p, _p]
_lib.rlipv2_box_pair_loss_f32.restype = _i
_u = ctypes.c_uint
_lib.rlipv2_short_attention_fwd_f32.argtypes = [_p, _p, _p, _p, _i, _i, _i, _i, _f, _d, _p, _u, _p, _p, _p]
_lib.rlipv2_short_attention_fwd_f32.restype = _i
_lib.rlipv2_short_attention_bwd_f32.argtypes = [_p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _d, _p, _u, _p, _p, _p, _p]
_lib.rlipv2_short_attention_bwd_f32.restype = _i
_lib.rlipv2_groupnorm_tokens_fwd_f32.argtypes = [_p, _p, _p, _f, _i, _i, _i, _i, _p, _p, _ll, _p, _p, _p]
_lib.rlipv2_groupnorm_tokens_fwd_f32.restype = _i
_lib.rlipv2_groupnorm_tokens_bwd_f32.argtypes = [_p, _ll, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p]
_lib.rlipv2_groupnorm_tokens_bwd_f32.restype = _i
for _n in ("wait_host_flag", "box_refine_f32", "sine_embed_f32"):
    getattr(_lib, "rlipv2_" + _n).restype = _i
for _n in ("add_layernorm_fwd_f32", "layernorm_bwd_f32", "relu_bwd_colsum_f32", "adamw_f32", "gather_chunks_f32", "rowmask_bwd_colsum_f32"):
    getattr(_lib, "rlipv2_" + _n).restype = _i
_lib.rlipv2_fused_error_string.argtypes = [_i]
_lib.rlipv2_fused_error_string.restype = ctypes.c_char_p
_lib.rlipv2_fused_launch_count.restype = ctypes.c_ulonglong

EXPORTS = ("rlipv2_add_layernorm_fwd_f32", "rlipv2_layernorm_bwd_f32", "rlipv2_relu_bwd_colsum_f32",
           "rlipv2_adamw_f32", "rlipv2_adamw_scaled_f32", "rlipv2_adamw_dev_f32", "rlipv2_gather_chunks_f32", "rlipv2_rowmask_bwd_colsum_f32", "rlipv2_wait_host_flag",
           "rlipv2_box_refine_f32", "rlipv2_sine_embed_f32", "rlipv2_box_pair_loss_f32", "rlipv2_groupnorm_tokens_fwd_f32", "rlipv2_groupnorm_tokens_bwd_f32",
           "rlipv2_short_attention_fwd_f32", "rlipv2_short_attention_bwd_f32", "rlipv2_stamp_globaltimer", "rlipv2_layernorm_bwd_acc_f32",
           "rlipv2_relu_bwd_colsum_acc_f32", "rlipv2_rowmask_bwd_colsum_acc_f32", "rlipv2_fused_error_string", "rlipv2_fused_launch_count")


def library_path():
    return _path


def launch_count():
    return int(_lib.rlipv2_fused_launch_count())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what}: {_lib.rlipv2_fused_error_string(rc).decode()} (code {rc})")


def ln_supported(C):
    return C % 128 == 0 and 0 < C <= 1024


def add_layernorm_fwd(x2, r2, gamma, beta, eps):
    """x2, r2 [M,C] contiguous fp32 CUDA (r2 may be None) -> (y, z, mean, rstd)"""
    M, C = x2.shape
    y = torch.empty_like(x2)
    z = torch.empty_like(x2)
    mean = torch.empty(M, dtype=torch.float32, device=x2.device)
    rstd = torch.empty(M, dtype=torch.float32, device=x2.device)
    with torch.cuda.device(x2.device):
        rc = _lib.rlipv2_add_layernorm_fwd_f32(x2.data_ptr(), r2.data_ptr() if r2 is not None else None, gamma.data_ptr(),
                                               beta.data_ptr(), eps, M, C, y.data_ptr(), z.data_ptr(), mean.data_ptr(),
                                               rstd.data_ptr(), _stream())
    _check(rc, "rlipv2_add_layernorm_fwd_f32")
    return y, z, mean, rstd


def _acc_ok(t, n):
    return t is not None and t.is_cuda and t.dtype == torch.float32 and t.numel() == n and t.is_contiguous()


def layernorm_bwd(dy2, z, mean, rstd, gamma, acc_gamma=None, acc_beta=None):
    """-> (dz, dgamma, dbeta).  With `acc_gamma` / `acc_beta` (contiguous fp32 [C] views of the flat gradient buffer)
    the parameter gradients are ADDED into them and (dz, None, None) is returned."""
    M, C = dy2.shape
    dz = torch.empty_like(dy2)
    acc = _acc_ok(acc_gamma, C) and _acc_ok(acc_beta, C)
    dgamma = acc_gamma if acc else torch.empty(C, dtype=torch.float32, device=dy2.device)
    dbeta = acc_beta if acc else torch.empty(C, dtype=torch.float32, device=dy2.device)
    with torch.cuda.device(dy2.device):
        rc = _lib.rlipv2_layernorm_bwd_acc_f32(dy2.data_ptr(), z.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                               gamma.data_ptr(), M, C, dz.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(),
                                               1 if acc else 0, _stream())
    _check(rc, "rlipv2_layernorm_bwd_acc_f32")
    return (dz, None, None) if acc else (dz, dgamma, dbeta)


def relu_bwd_colsum(g2, y2=None, acc=None):
    """g2 [M,N] contiguous; y2 = forward output of the ReLU (or None) -> (masked g [M,N], column sums [N]).
    With `acc` (contiguous fp32 [N] view of the flat gradient buffer) the column sums are ADDED into it and
    (masked g, None) is returned."""
    M, N = g2.shape
    into = _acc_ok(acc, N)
    colsum = acc if into else torch.empty(N, dtype=torch.float32, device=g2.device)
    gm = torch.empty_like(g2) if y2 is not None else g2
    with torch.cuda.device(g2.device):
        rc = _lib.rlipv2_relu_bwd_colsum_acc_f32(g2.data_ptr(), y2.data_ptr() if y2 is not None else None,
                                                 gm.data_ptr() if y2 is not None else None, colsum.data_ptr(), M, N,
                                                 1 if into else 0, _stream())
    _check(rc, "rlipv2_relu_bwd_colsum_acc_f32")
    return gm, (None if into else colsum)


def adamw(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step, grad_scale=None, lr_dev=None,
          skip_flag=None):
    """In-place AdamW over flat contiguous fp32 buffers; `step` = 0-dim device float (1-based count); `grad_scale` =
    optional 0-dim / 1-element device float the gradient is multiplied by as it is read (clip coefficient, 1 / world);
    `lr_dev` = optional 1-element device float that overrides `lr` (a captured graph then follows the scheduler);
    `skip_flag` = optional 1-element device int32: non-zero makes the update a no-op."""
    with torch.cuda.device(param.device):
        rc = _lib.rlipv2_adamw_dev_f32(param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
                                       param.numel(), lr, lr_dev.data_ptr() if lr_dev is not None else None, beta1, beta2,
                                       eps, weight_decay, step.data_ptr(),
                                       grad_scale.data_ptr() if grad_scale is not None else None,
                                       skip_flag.data_ptr() if skip_flag is not None else None, _stream())
    _check(rc, "rlipv2_adamw_dev_f32")


GATHER_CHUNK = 32768


def gather_table(tensors, offsets, out=None):
    """Host side of `gather_chunks`: int64 [n_chunks, 3] rows (source address, destination element offset, count)
    that copy tensors[i] (contiguous fp32 CUDA) to flat[offsets[i] : offsets[i] + numel].  `out`: optional
    (pinned) int64 tensor to fill; returns (table, n_chunks)."""
    rows = []
    for t, off in zip(tensors, offsets):
        base, n = t.data_ptr(), t.numel()
        for s in range(0, n, GATHER_CHUNK):
            rows.append((base + 4 * s, off + s, min(GATHER_CHUNK, n - s)))
    tab = torch.tensor(rows, dtype=torch.int64).reshape(-1, 3)
    if out is not None:
        out[:tab.shape[0]].copy_(tab)
        return out, tab.shape[0]
    return tab, tab.shape[0]


def gather_chunks(table_dev, n_chunks, flat):
    """flat (contiguous fp32 CUDA) <- chunks described by `table_dev` (int64 [>= n_chunks, 3] on the device)"""
    with torch.cuda.device(flat.device):
        rc = _lib.rlipv2_gather_chunks_f32(table_dev.data_ptr(), n_chunks, flat.data_ptr(), _stream())
    _check(rc, "rlipv2_gather_chunks_f32")


def rowmask_bwd_colsum(g2, rowmask, acc=None):
    """g2 [M,N] contiguous fp32, rowmask [M] bool (True = row zeroed in the forward) -> (masked g, column sums);
    `acc`: as in relu_bwd_colsum"""
    M, N = g2.shape
    into = _acc_ok(acc, N)
    colsum = acc if into else torch.empty(N, dtype=torch.float32, device=g2.device)
    gm = torch.empty_like(g2)
    with torch.cuda.device(g2.device):
        rc = _lib.rlipv2_rowmask_bwd_colsum_acc_f32(g2.data_ptr(), rowmask.data_ptr(), gm.data_ptr(), colsum.data_ptr(), M, N,
                                                    1 if into else 0, _stream())
    _check(rc, "rlipv2_rowmask_bwd_colsum_acc_f32")
    return gm, (None if into else colsum)


def wait_host_flag(flag_pinned, seq_dev, err_dev, timeout_s=10.0):
    """Enqueue the backward graph's head: hold the current stream until the pinned int32 word `flag_pinned[0]`
    reaches seq_dev[0] + 1 (see include/rlipv2_fused.h).  flag_pinned: pinned CPU int32 tensor (UVA: its host
    address is its device address); seq_dev, err_dev: CUDA int32 tensors."""
    assert flag_pinned.is_pinned() and flag_pinned.dtype == torch.int32
    assert seq_dev.is_cuda and err_dev.is_cuda and seq_dev.dtype == torch.int32 and err_dev.dtype == torch.int32
    with torch.cuda.device(seq_dev.device):
        rc = _lib.rlipv2_wait_host_flag(flag_pinned.data_ptr(), seq_dev.data_ptr(), int(timeout_s * 1e9),
                                        err_dev.data_ptr(), _stream())
    _check(rc, "rlipv2_wait_host_flag")


def box_refine(delta, ref, eps=1e-5):
    """sigmoid(delta + inverse_sigmoid(ref)); contiguous fp32 CUDA tensors of equal shape."""
    y = torch.empty_like(delta)
    with torch.cuda.device(delta.device):
        rc = _lib.rlipv2_box_refine_f32(delta.data_ptr(), ref.data_ptr(), eps, delta.numel(), y.data_ptr(), _stream())
    _check(rc, "rlipv2_box_refine_f32")
    return y


def box_pair_loss(src, tgt):
    """src, tgt [R, 4] contiguous fp32 CUDA (cx, cy, w, h) -> l1 [R], giou_loss [R], dl1 [R, 4], dgiou [R, 4]"""
    R = src.shape[0]
    l1 = torch.empty(R, dtype=torch.float32, device=src.device)
    gl = torch.empty_like(l1)
    dl1, dgl = torch.empty_like(src), torch.empty_like(src)
    with torch.cuda.device(src.device):
        rc = _lib.rlipv2_box_pair_loss_f32(src.data_ptr(), tgt.data_ptr(), R, l1.data_ptr(), gl.data_ptr(),
                                           dl1.data_ptr(), dgl.data_ptr(), _stream())
    _check(rc, "rlipv2_box_pair_loss_f32")
    return l1, gl, dl1, dgl


def groupnorm_tokens_fwd(x_tok, gamma, beta, eps, out_rows, out_batch_stride, groups=32):
    """x_tok [N, HW, C] contiguous fp32 CUDA -> writes GroupNorm(x) into `out_rows` (a view whose data_ptr is the
    level's first row of a [N, S, C] buffer; rows of image n start at n * out_batch_stride elements).
    -> (mean [N, G], rstd [N, G])"""
    N, HW, C = x_tok.shape
    stats = torch.empty((N, groups, 2), dtype=torch.float64, device=x_tok.device)
    mean = torch.empty((N, groups), dtype=torch.float32, device=x_tok.device)
    rstd = torch.empty_like(mean)
    with torch.cuda.device(x_tok.device):
        rc = _lib.rlipv2_groupnorm_tokens_fwd_f32(x_tok.data_ptr(), gamma.data_ptr(), beta.data_ptr(), eps, N, HW, C, groups,
                                                  stats.data_ptr(), out_rows.data_ptr(), out_batch_stride, mean.data_ptr(),
                                                  rstd.data_ptr(), _stream())
    _check(rc, "rlipv2_groupnorm_tokens_fwd_f32")
    return mean, rstd


def groupnorm_tokens_bwd(dy_rows, dy_batch_stride, x_tok, mean, rstd, gamma, dgamma, dbeta, groups=32):
    """dy_rows: view at the level's first row of the [N, S, C] output gradient; dgamma / dbeta [C] are added to.
    -> dx [N, HW, C]"""
    N, HW, C = x_tok.shape
    sums = torch.empty((N, groups, 2), dtype=torch.float64, device=x_tok.device)
    dx = torch.empty_like(x_tok)
    with torch.cuda.device(x_tok.device):
        rc = _lib.rlipv2_groupnorm_tokens_bwd_f32(dy_rows.data_ptr(), dy_batch_stride, x_tok.data_ptr(), mean.data_ptr(),
                                                  rstd.data_ptr(), gamma.data_ptr(), N, HW, C, groups, sums.data_ptr(),
                                                  dx.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), _stream())
    _check(rc, "rlipv2_groupnorm_tokens_bwd_f32")
    return dx


SHORT_ATTN_MAX_T, SHORT_ATTN_D = 8, 64


def short_attention_fwd(q, k, v, mask, scale, dropout_p, seed, salt):
    """q, k, v [B, T, H, 64] contiguous fp32 CUDA (T <= 8); mask additive [B, T] or None; seed: device int64 [1] (needed
    when dropout_p > 0) -> (out [B, T, H, 64], seed_used int64 [1])"""
    B, T, H, D = q.shape
    out = torch.empty_like(q)
    seed_used = torch.empty(1, dtype=torch.int64, device=q.device)
    with torch.cuda.device(q.device):
        rc = _lib.rlipv2_short_attention_fwd_f32(q.data_ptr(), k.data_ptr(), v.data_ptr(),
                                                 mask.data_ptr() if mask is not None else None, B, H, T, D, scale,
                                                 float(dropout_p), seed.data_ptr() if seed is not None else None, int(salt),
                                                 out.data_ptr(), seed_used.data_ptr(), _stream())
    _check(rc, "rlipv2_short_attention_fwd_f32")
    return out, seed_used


def short_attention_bwd(q, k, v, mask, grad_out, scale, dropout_p, seed_used, salt):
    B, T, H, D = q.shape
    dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
    with torch.cuda.device(q.device):
        rc = _lib.rlipv2_short_attention_bwd_f32(q.data_ptr(), k.data_ptr(), v.data_ptr(),
                                                 mask.data_ptr() if mask is not None else None, grad_out.data_ptr(), B, H, T, D,
                                                 scale, float(dropout_p), seed_used.data_ptr(), int(salt), dq.data_ptr(),
                                                 dk.data_ptr(), dv.data_ptr(), _stream())
    _check(rc, "rlipv2_short_attention_bwd_f32")
    return dq, dk, dv


def sine_embed(pos2d):
    """pos2d [R, n] contiguous fp32 CUDA, n in (2, 4) -> [R, n*128]"""
    R, n = pos2d.shape
    out = torch.empty((R, n * 128), dtype=torch.float32, device=pos2d.device)
    with torch.cuda.device(pos2d.device):
        rc = _lib.rlipv2_sine_embed_f32(pos2d.data_ptr(), R, n, out.data_ptr(), _stream())
    _check(rc, "rlipv2_sine_embed_f32")
    return out


def stamp(dst_int64, index):
    """diagnostic: dst_int64[index] = GPU globaltimer (ns) when the current stream gets here"""
    assert dst_int64.is_cuda and dst_int64.dtype == torch.int64
    with torch.cuda.device(dst_int64.device):
        rc = _lib.rlipv2_stamp_globaltimer(dst_int64.data_ptr() + 8 * index, _stream())
    _check(rc, "rlipv2_stamp_globaltimer")
