"""Dense-contraction seam of the hot path.

Every GEMM-shaped op of ALIF, the RobertaLayer stack, the text tower and the deformable encoder/decoder goes through
these functions, so the module code stays independent of how the contraction is executed.

Three execution modes (`set_matmul_precision`):
  'fp32'  IEEE fp32 products through cuBLAS - used by the parity tests against the reference fixtures;
  '3xtf32' the tcgen05 kernels with error-compensated operands: x = hi + lo with `hi` the TF32
          truncation, x w ~ hi(x) hi(w) + hi(x) lo(w) + lo(x) hi(w) evaluated as ONE tcgen05 GEMM over the three-fold
          contraction axis.  fp32-class accuracy (error ~2^-20), so the 1e-3 model-level parity tests run through the same
          linear / weight-gradient / input-gradient kernels the benchmark uses; everything else as in 'fp32';
  'tf32'  TF32 tensor-core products, fp32 accumulation - what the reference's pinned torch 1.10 does by default on
          tensor-core GPUs, and what bench.py measures.  Forward linears whose shape the hand-written tcgen05 kernel
          supports (N % 128 == 0, K % 32 == 0; csrc/dense_tf32.cu, include/rlipv2_dense.h) run on it with bias / ReLU /
          GELU / row mask fused in the epilogue.  Backward: split-K tcgen05 weight gradients for tall-skinny shapes, the
          encoder FFN's gated input gradient + both weight gradients on tcgen05 (`_FFNReLU`), cuBLAS TF32 for the rest;
          bias / LayerNorm gradients from the fused kernels of csrc/fused_ops.cu.
What the custom autograd functions add on top of the kernels (all of it only when the train step has made the parameters'
`.grad` views of its flat gradient buffer and marked them `_fuse_grad`):
  * parameter gradients are ADDED into those views by the producing kernel (GEMM beta = 1, reductions without zero-fill)
    instead of being handed to ~600 AccumulateGrad kernels;
  * they are issued on a side stream (`_ParamGradSide`), off the chain of input gradients the previous layer waits for;
    `join_param_grad_stream()` must be called before the gradients are read.
There is no CPU implementation behind the tcgen05 path; CPU tensors only ever reach the torch ops.
"""
import math
import os

import torch
import torch.nn.functional as F

from . import streams

_PRECISION = "fp32"
_USE_TCGEN05 = True
_USE_FUSED = True        # fused LayerNorm / bias-gradient kernels (exact fp32 arithmetic; CUDA tensors only)
# backward GEMMs (weight / input gradients) on the tcgen05 kernels instead of cuBLAS, for calls with at least
# _OWN_BWD_MIN_ROWS rows (the split-K weight gradient pays off on the encoder's 44k-token activations)
_OWN_BWD = os.environ.get("RLIPV2_OWN_BWD", "0") == "1"
# 'hybrid' FFN backward: only the down-projection's input gradient (ReLU gate + bias gradient fused in its epilogue) and
# the two split-K weight gradients run on the tcgen05 kernels; the plain input gradient stays on cuBLAS' 256x256 2-SM kernel
# (measured on B200, r01s4a: 32.79 vs 33.40 ms/step -> the default; RLIPV2_FFN_BWD=cublas restores the torch backward)
_FFN_BWD = os.environ.get("RLIPV2_FFN_BWD", "hybrid")
_OWN_WGRAD = os.environ.get("RLIPV2_OWN_WGRAD", "1") != "0"
_OWN_BWD_MIN_ROWS = int(os.environ.get("RLIPV2_OWN_BWD_MIN_ROWS", "4096"))
# Parameter gradients on a side stream.  Only the input gradient of a layer is on the chain to the previous layer; the
# weight / bias gradient is not needed before the optimizer.  For the small linears (decoders, heads, ALIF, RobertaLayer,
# text tower: chains of launch-bound 5-8 us GEMMs) this halves the chain; for the encoder's 44k-token layers the weight
# gradients fill the SMs the latency-bound MSDeformAttn backward leaves idle.  Active only when the gradient is
# accumulated in place into the step's flat gradient buffer (`_fuse_grad`), so nothing is handed back to autograd from
# the side stream; the train step joins it after backward() (`join_param_grad_stream`).
_LIBRARY_SMALL = os.environ.get("RLIPV2_TEXT_LIBRARY_GEMM", "0") != "0"      # measured r01s4g: 27.6 vs 27.7 ms/step - no gain, off
_LIBRARY_SMALL_MAX_ROWS = 4096
# split-K tcgen05 forward for small-M / long-K linears (ALIF out projections, label-side in-projections, RobertaLayer
# FFN-down): measured r02a 27.19 vs 27.41 ms/step on one box, tests/test_zz5_splitk_gpu.py green -> on by default
_SPLITK_FWD = os.environ.get("RLIPV2_SPLITK_FWD", "1") != "0"
_WGRAD_STREAM = os.environ.get("RLIPV2_WGRAD_STREAM", "1") != "0"
# (measured r01s4d: every size on the side stream 28.65 vs 29.5 ms/step with only the <= 4096-row problems there)
_WGRAD_STREAM_MAX_ROWS = int(os.environ.get("RLIPV2_WGRAD_STREAM_MAX_ROWS", str(1 << 30)))
_param_grad_streams = {}
_param_grad_pending = set()


def _param_grad_stream(device):
    st = _param_grad_streams.get(device)
    if st is None:
        st = _param_grad_streams[device] = streams.get(device, "param_grad")
    return st


class _ParamGradSide:
    """`with _ParamGradSide(dev, t1, t2, ...)`: the body runs on the parameter-gradient side stream, after everything
    issued so far on the current stream; the tensors it reads are kept alive for that stream."""

    def __init__(self, device, *reads):
        self.device, self.reads = device, reads

    def __enter__(self):
        self.side = _param_grad_stream(self.device)
        self.side.wait_stream(torch.cuda.current_stream(self.device))
        self.ctx = torch.cuda.stream(self.side)
        self.ctx.__enter__()
        return self

    def __exit__(self, *a):
        self.ctx.__exit__(*a)
        for t in self.reads:
            if t is not None:
                t.record_stream(self.side)
        _param_grad_pending.add(self.device)


def pending_param_grad_streams():
    """the side streams that hold parameter-gradient work issued since the last join"""
    return [_param_grad_streams[dev] for dev in _param_grad_pending]


def join_param_grad_stream(device=None):
    """make the current stream wait for the parameter-gradient side stream(s) (call after backward(), before the
    gradients are read)"""
    for dev in list(_param_grad_pending):
        if device is None or dev == device:
            torch.cuda.current_stream(dev).wait_stream(_param_grad_streams[dev])
            _param_grad_pending.discard(dev)


def _split_tf32(t):
    """t = hi + lo, hi = t with the 13 low mantissa bits cleared (exactly what a TF32 tensor-core operand keeps)"""
    hi = (t.view(torch.int32) & -8192).view(torch.float32)
    return hi, t - hi


def _linear_kernel(x2, w, bias, act, rm):
    """act(x2 w^T + bias) on the tcgen05 kernel; '3xtf32': one GEMM over [hi | hi | lo] x [hi | lo | hi]"""
    abi = _abi()
    if _PRECISION == "3xtf32":
        xh, xl = _split_tf32(x2)
        wh, wl = _split_tf32(w)
        return abi.linear_tf32(torch.cat((xh, xh, xl), 1), torch.cat((wh, wl, wh), 1), bias, act, rm)
    return abi.linear_tf32(x2, w, bias, act, rm)


def _wgrad_kernel(g, x2, acc=None):
    """dw[N,K] (+)= g[T,N]^T x2[T,K] on the split-K tcgen05 kernel (contraction over the rows)"""
    if _PRECISION == "3xtf32":
        gh, gl = _split_tf32(g)
        xh, xl = _split_tf32(x2)
        return _abi().wgrad_tf32(torch.cat((gh, gh, gl), 0), torch.cat((xh, xl, xh), 0), acc=acc)
    return _abi().wgrad_tf32(g, x2, acc=acc)


def _dgrad_kernel(g, w):
    """dx[T,K] = g[T,N] w[N,K] on the tcgen05 kernel (contraction over N)"""
    if _PRECISION == "3xtf32":
        gh, gl = _split_tf32(g)
        wh, wl = _split_tf32(w)
        return _abi().dgrad_tf32(torch.cat((gh, gh, gl), 1), torch.cat((wh, wl, wh), 0))[0]
    return _abi().dgrad_tf32(g, w)[0]


def set_matmul_precision(mode: str, tcgen05: bool = True, fused: bool = True):
    global _PRECISION, _USE_TCGEN05, _USE_FUSED
    assert mode in ("fp32", "tf32", "3xtf32")
    _PRECISION = mode
    _USE_TCGEN05 = tcgen05
    _USE_FUSED = fused
    torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
    torch.backends.cudnn.allow_tf32 = mode == "tf32"


def matmul_precision():
    return _PRECISION


def _abi():
    from . import dense_abi
    return dense_abi


class _LinearTF32(torch.autograd.Function):
    """y = act(x W^T + b) on the tcgen05 kernel; backward = cuBLAS TF32 GEMMs + fused mask."""

    @staticmethod
    def forward(ctx, x, weight, bias, act, row_mask=None, library_small=False):
        abi = _abi()
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        w = weight if weight.is_contiguous() else weight.contiguous()
        rm = None
        if row_mask is not None:
            rm = row_mask.reshape(-1)
            rm = rm if rm.is_contiguous() else rm.contiguous()
        if (library_small and act == 0 and rm is None and x2.shape[0] <= _LIBRARY_SMALL_MAX_ROWS and w.shape[1] >= 512
                and _PRECISION == "tf32"):
            # plain small-M / long-K linears of the (third-party, HF) text tower: cuBLAS' split-K kernels are 2-3x faster
            # than one 128-row tile per CTA here (profiles/dense_microbench_r01_v3_small_grids.jsonl); backward unchanged
            y = F.linear(x2, w, bias)
        elif _PRECISION == "3xtf32":
            y = _linear_kernel(x2, w, bias, act, rm)
        elif _SPLITK_FWD and act == 0 and rm is None and abi.splitk_splits(x2.shape[0], w.shape[0], w.shape[1]) > 1:
            # small-M / long-K (ALIF out projections, label-side in-projections, RobertaLayer FFN-down): K split over the
            # SMs one-CTA-per-tile leaves idle (rlipv2_dense_linear_splitk_tf32); backward unchanged
            y = abi.linear_splitk_tf32(x2, w, bias, abi.splitk_splits(x2.shape[0], w.shape[0], w.shape[1]))
        else:
            y = _linear_kernel(x2, w, bias, act, rm)
        ctx.act = act
        ctx.has_bias = bias is not None
        # parameters whose .grad is a view of the step's flat gradient buffer (train_step marks them `_fuse_grad`):
        # the backward adds their gradients straight into that view instead of returning them to AccumulateGrad
        ctx.w_acc = weight if getattr(weight, "_fuse_grad", False) and weight.is_contiguous() else None
        ctx.b_acc = bias if bias is not None and getattr(bias, "_fuse_grad", False) else None
        ctx.save_for_backward(x2, w, y if act == abi.ACT_RELU else None, rm)
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, grad_out):
        x2, w, y, rm = ctx.saved_tensors
        g = grad_out.reshape(-1, grad_out.shape[-1])
        if not g.is_contiguous():
            g = g.contiguous()
        gb = None
        b_acc = ctx.b_acc.grad if ctx.b_acc is not None and ctx.needs_input_grad[2] else None
        w_acc = ctx.w_acc.grad if ctx.w_acc is not None and ctx.needs_input_grad[1] else None
        T, N, K = g.shape[0], g.shape[1], w.shape[1]
        # parameter gradients off the critical chain (small problems whose gradients accumulate in place)
        side = None
        if (_WGRAD_STREAM and g.is_cuda and T <= _WGRAD_STREAM_MAX_ROWS and w_acc is not None
                and (not ctx.has_bias or not ctx.needs_input_grad[2] or b_acc is not None)):
            side = _param_grad_stream(g.device)
        plain_bias = rm is None and y is None and g.shape[1] % 32 == 0           # column sums only, g unchanged
        if rm is not None:
            # rows zeroed in the forward carry no gradient (N % 128 == 0 on this path, so the kernel applies)
            g, gb = _fused().rowmask_bwd_colsum(g, rm, acc=b_acc)
        elif g.shape[1] % 32 == 0:
            if not (side is not None and plain_bias):
                # one pass: ReLU mask (when fused in the forward) + bias-gradient column sum
                g, gb = _fused().relu_bwd_colsum(g, y, acc=b_acc)
        else:
            if y is not None:
                g = g * (y > 0)
            gb = g.sum(0)
            if side is not None and b_acc is not None:
                b_acc.add_(gb)
                gb = None
        gx = gw = None
        big = T >= _OWN_BWD_MIN_ROWS and _abi().grads_supported(T, N, K)
        if side is not None:
            with _ParamGradSide(g.device, g, x2):        # g (masked) is complete on the current stream; x2 long since
                if plain_bias and ctx.has_bias and ctx.needs_input_grad[2]:
                    _fused().relu_bwd_colsum(g, None, acc=b_acc)
                if big and _OWN_WGRAD and (_OWN_BWD or N * K <= 384 * 256):
                    _wgrad_kernel(g, x2, acc=w_acc)      # split-K tcgen05 kernel, reduces straight into the view
                else:
                    w_acc.addmm_(g.t(), x2)              # cuBLAS with beta = 1: grad view += g^T x
            if ctx.needs_input_grad[0]:
                gx = (_dgrad_kernel(g, w) if (_OWN_BWD or _PRECISION == "3xtf32") and big else g @ w).view(*grad_out.shape[:-1], K)
            return gx, None, None, None, None, None
        if ctx.needs_input_grad[0]:
            gx = (_dgrad_kernel(g, w) if (_OWN_BWD or _PRECISION == "3xtf32") and big else g @ w).view(*grad_out.shape[:-1], K)
        if ctx.needs_input_grad[1]:
            # tall-skinny weight gradients (44k tokens -> a 256x256 .. 384x256 weight): cuBLAS falls back to an
            # sm_80 64x64 kernel at 65-72 us; the split-K tcgen05 kernel takes 28-40 us (measured, B200)
            if big and _OWN_WGRAD and (_OWN_BWD or N * K <= 384 * 256):
                gw = _wgrad_kernel(g, x2, acc=w_acc)
                gw = None if w_acc is not None else gw
            elif w_acc is not None:
                w_acc.addmm_(g.t(), x2)                  # cuBLAS with beta = 1: grad view += g^T x
                gw = None
            else:
                gw = g.t() @ x2
        if not (ctx.has_bias and ctx.needs_input_grad[2]):
            gb = None
        return gx, gw, gb, None, None, None


class _FFNReLU(torch.autograd.Function):
    """y = relu(x W1^T + b1) W2^T + b2 - the position-wise FFN of the deformable encoder / decoder layers
    (dab_deformable/deformable_transformer.py:1283-1287, 1368-1372; dropout p = 0 in every ParSeDA script).
    Forward: two tcgen05 GEMMs (bias + ReLU fused).  Backward, all on the tcgen05 kernels of csrc/dense_tf32.cu:
        dW2 = g^T h            split-K weight gradient, operands read as stored
        gh  = (g W2) * (h > 0) input gradient with the ReLU mask and db1 = colsum(gh) fused in the epilogue
        dW1 = gh^T x, dx = gh W1,  db2 = colsum(g)
    which removes the separate mask + column-sum pass over the [T, d_ffn] hidden gradient (1.1 GB of HBM
    traffic per encoder layer at 800x1333, batch 2)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        abi = _abi()
        x2 = x.reshape(-1, x.shape[-1])
        x2 = x2 if x2.is_contiguous() else x2.contiguous()
        h = abi.linear_tf32(x2, w1.contiguous(), b1, abi.ACT_RELU)
        y = abi.linear_tf32(h, w2.contiguous(), b2, abi.ACT_NONE)
        ctx.save_for_backward(x2, w1, w2, h)
        ctx.params = (w1, b1, w2, b2)
        return y.view(*x.shape[:-1], w2.shape[0])

    @staticmethod
    def backward(ctx, grad_out):
        abi = _abi()
        x2, w1, w2, h = ctx.saved_tensors
        g = grad_out.reshape(-1, grad_out.shape[-1])
        g = g if g.is_contiguous() else g.contiguous()
        acc = lambda p: p.grad if (p is not None and getattr(p, "_fuse_grad", False) and p.grad is not None) else None
        w1p, b1p, w2p, b2p = ctx.params
        T = g.shape[0]
        if (_WGRAD_STREAM and T <= _WGRAD_STREAM_MAX_ROWS and all(acc(p) is not None for p in ctx.params)):
            # the chain to the previous layer is gh -> gx; the two weight gradients and db2 run beside it
            with _ParamGradSide(g.device, g, h):
                _fused().relu_bwd_colsum(g, None, acc=acc(b2p))
                abi.wgrad_tf32(g, h, acc=acc(w2p))
            gh, gb1 = abi.dgrad_tf32(g, w2, relu_out=h)
            b1p.grad.add_(gb1)
            with _ParamGradSide(g.device, gh, x2):
                abi.wgrad_tf32(gh, x2, acc=acc(w1p))
            gx = None
            if ctx.needs_input_grad[0]:
                gx = (gh @ w1 if _FFN_BWD == "hybrid" else abi.dgrad_tf32(gh, w1)[0]).view(*grad_out.shape[:-1], w1.shape[1])
            return gx, None, None, None, None
        _, gb2 = _fused().relu_bwd_colsum(g, None, acc=acc(b2p))
        gw2 = abi.wgrad_tf32(g, h, acc=acc(w2p))
        gh, gb1 = abi.dgrad_tf32(g, w2, relu_out=h)
        gw1 = abi.wgrad_tf32(gh, x2, acc=acc(w1p))
        gx = None
        if ctx.needs_input_grad[0]:
            gx = (gh @ w1 if _FFN_BWD == "hybrid" else abi.dgrad_tf32(gh, w1)[0]).view(*grad_out.shape[:-1], w1.shape[1])
        if acc(b1p) is not None:
            b1p.grad.add_(gb1)
            gb1 = None
        gw1 = None if acc(w1p) is not None else gw1
        gw2 = None if acc(w2p) is not None else gw2
        return gx, gw1, gb1, gw2, gb2


def ffn_relu(x, w1, b1, w2, b2):
    """relu(x W1^T + b1) W2^T + b2"""
    M = x.numel() // x.shape[-1]
    if (_PRECISION == "tf32" and (_OWN_BWD or _FFN_BWD == "hybrid") and _tcgen05_ok(x, w1) and _abi().supported(M, w2.shape[0], w2.shape[1]) and M >= _OWN_BWD_MIN_ROWS
            and b1 is not None and b2 is not None and w2.shape[0] % 32 == 0):
        return _FFNReLU.apply(x, w1, b1, w2, b2)
    return linear(linear_relu(x, w1, b1), w2, b2)


def _fused():
    from . import fused_abi
    return fused_abi


class _AddLayerNorm(torch.autograd.Function):
    """y = LayerNorm(x + r) (r optional) with the fused forward / backward kernels of csrc/fused_ops.cu."""

    @staticmethod
    def forward(ctx, x, r, weight, bias, eps):
        f = _fused()
        C = x.shape[-1]
        x2 = x.reshape(-1, C)
        x2 = x2 if x2.is_contiguous() else x2.contiguous()
        r2 = None
        if r is not None:
            r2 = r.reshape(-1, C)
            r2 = r2 if r2.is_contiguous() else r2.contiguous()
        y, z, mean, rstd = f.add_layernorm_fwd(x2, r2, weight, bias, eps)
        ctx.save_for_backward(z, mean, rstd, weight)
        ctx.has_r = r is not None
        fuse = getattr(weight, "_fuse_grad", False) and getattr(bias, "_fuse_grad", False)
        ctx.acc = (weight, bias) if fuse else None
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, grad_out):
        z, mean, rstd, weight = ctx.saved_tensors
        g = grad_out.reshape(-1, grad_out.shape[-1])
        g = g if g.is_contiguous() else g.contiguous()
        acc = ctx.acc if ctx.acc is not None and ctx.needs_input_grad[2] and ctx.needs_input_grad[3] else None
        if acc is not None and acc[0].grad is not None and acc[1].grad is not None:
            dz, dgamma, dbeta = _fused().layernorm_bwd(g, z, mean, rstd, weight, acc[0].grad, acc[1].grad)
        else:
            dz, dgamma, dbeta = _fused().layernorm_bwd(g, z, mean, rstd, weight)
        dz = dz.view(grad_out.shape)
        return dz, (dz if ctx.has_r else None), dgamma, dbeta, None


def _fused_ln_ok(x):
    return _USE_FUSED and x.is_cuda and x.dtype == torch.float32 and _fused().ln_supported(x.shape[-1]) and x.numel() > 0


def _tcgen05_ok(x, weight):
    if not (_USE_TCGEN05 and _PRECISION in ("tf32", "3xtf32") and x.is_cuda and x.dtype == torch.float32):
        return False
    M = x.numel() // x.shape[-1]
    return M > 0 and _abi().supported(M, weight.shape[0], weight.shape[1])


def linear(x, weight, bias=None, row_mask=None, library_small=False):
    """x W^T + b; `row_mask` (bool, x.shape[:-1]): those rows of the result are zero (masked_fill folded in);
    `library_small`: the caller allows the cuBLAS forward for small-M / long-K problems (text tower)"""
    if _tcgen05_ok(x, weight):
        return _LinearTF32.apply(x, weight, bias, 0, row_mask, library_small and _LIBRARY_SMALL)
    y = F.linear(x, weight, bias)
    return y if row_mask is None else y.masked_fill(row_mask[..., None], float(0))


def linear_relu(x, weight, bias=None):
    if _tcgen05_ok(x, weight):
        return _LinearTF32.apply(x, weight, bias, 1)
    return F.relu(F.linear(x, weight, bias))


def linear_gelu(x, weight, bias=None):
    if _tcgen05_ok(x, weight):
        if not (torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad)):
            return _LinearTF32.apply(x, weight, bias, 2)
        return F.gelu(_LinearTF32.apply(x, weight, bias, 0))     # GELU backward needs the pre-activation
    return F.gelu(F.linear(x, weight, bias))


def layer_norm(x, weight, bias, eps):
    if _fused_ln_ok(x):
        return _AddLayerNorm.apply(x, None, weight, bias, eps)
    return F.layer_norm(x, (x.shape[-1],), weight, bias, eps)


def add_layer_norm(x, residual, weight, bias, eps):
    """LayerNorm(x + residual)."""
    if _fused_ln_ok(x) and residual.shape == x.shape:
        return _AddLayerNorm.apply(x, residual, weight, bias, eps)
    return F.layer_norm(x + residual, (x.shape[-1],), weight, bias, eps)


# ---- softmax-attention cores (csrc/attn_tf32.cu, include/rlipv2_attn.h) ---------------------------------------------
# ALIF's bidirectional cross-attention (fuse_helper.py:395-445), the RobertaLayer's 12 x 64 self-attention
# (modeling_roberta.py:185-241) and the decoders' query self-attention (dab_deformable/deformable_transformer.py:1383-1390)
# are the same bmm -> softmax -> dropout -> bmm chain; in 'tf32' mode on a GPU it is ONE fused tcgen05 kernel forward and
# one dS kernel + three batched tcgen05 GEMMs backward, on [B, T, H*D] tensors as the projections produce them.
_ATTN = os.environ.get("RLIPV2_FUSED_ATTN", "1") != "0"
_dropout_seeds = {}


def dropout_seed(device):
    """device-resident int64 counter the fused attention kernels hash their dropout masks from"""
    s = _dropout_seeds.get(device)
    if s is None:
        s = _dropout_seeds[device] = torch.zeros(1, dtype=torch.int64, device=device) + (torch.initial_seed() % (2 ** 62))
    return s


def advance_dropout_seed(device):
    """fresh dropout masks for the next forward (one tiny kernel; safe inside a CUDA-graph capture)"""
    dropout_seed(device).add_(1)


def _attn_abi():
    from . import attn_abi
    return attn_abi


class _Attention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, key_bias, heads, scale, dropout_p, salt):
        abi = _attn_abi()
        q_, k_, v_ = (t if abi.usable(t) else t.contiguous() for t in (q, k, v))
        kb = None
        if key_bias is not None:
            kb = key_bias if key_bias.is_contiguous() else key_bias.contiguous()
        seed = dropout_seed(q.device) if dropout_p > 0 else None
        out, stats, seed_used = abi.forward(q_, k_, v_, heads, kb, scale, dropout_p, seed, salt)
        ctx.save_for_backward(q_, k_, v_, kb, out, stats, seed_used)
        ctx.conf = (heads, scale, dropout_p, salt)
        return out

    @staticmethod
    def backward(ctx, dout):
        q_, k_, v_, kb, out, stats, seed_used = ctx.saved_tensors
        heads, scale, dropout_p, salt = ctx.conf
        g = dout if dout.is_contiguous() else dout.contiguous()
        dq, dk, dv = _attn_abi().backward(q_, k_, v_, heads, kb, out, g, stats, scale, dropout_p, seed_used, salt)
        return dq, dk, dv, None, None, None, None, None


def _attn_ok(q, k, v, heads):
    if not (_ATTN and _USE_TCGEN05 and _PRECISION == "tf32" and q.is_cuda and q.dtype == torch.float32 and q.dim() == 3):
        return False
    B, Tq, C = q.shape
    return (k.shape[0] == B and v.shape == k.shape and k.shape[2] == C and C % heads == 0
            and _attn_abi().supported(B, heads, Tq, k.shape[1], C // heads))


def attention(q, k, v, heads, scale, key_bias=None, dropout_p=0.0, training=False, salt=0):
    """softmax_j(scale * <q_i, k_j> + key_bias_j) -> dropout -> . v, per head.
    q [B, Tq, H*D], k / v [B, Nk, H*D] (head h = columns [h*D, (h+1)*D)), key_bias [B, Nk] additive or None
    -> [B, Tq, H*D]"""
    p = float(dropout_p) if training else 0.0
    if _attn_ok(q, k, v, heads):
        return _Attention.apply(q, k, v, key_bias, heads, float(scale), p, int(salt))
    # torch path: CPU host-logic tests, the IEEE-fp32 parity mode and problems beyond the kernel's 480 keys - the
    # reference's own op chain
    B, Tq, C = q.shape
    D = C // heads
    qh = q.reshape(B, Tq, heads, D).transpose(1, 2)
    kh = k.reshape(B, -1, heads, D).transpose(1, 2)
    vh = v.reshape(B, -1, heads, D).transpose(1, 2)
    scores = torch.matmul(qh * scale, kh.transpose(-1, -2))
    if key_bias is not None:
        scores = scores + key_bias[:, None, None, :]
    probs = torch.softmax(scores, dim=-1)
    if p > 0:
        probs = F.dropout(probs, p=p, training=True)
    return torch.matmul(probs, vh).transpose(1, 2).reshape(B, Tq, C)


def bi_attention(q, k, vv, vl, heads, scale, dropout_p=0.0, training=False, salt=0):
    """ALIF's two directions over one score matrix S = scale * q k^T (fuse_helper.py:395-445):
    out_v = dropout(softmax_rows(S)) vl  (image tokens attend to labels),
    out_l = dropout(softmax_rows(S^T)) vv (labels attend to image tokens) - i.e. attention(q, k, vl) and attention(k, q, vv)."""
    out_v = attention(q, k, vl, heads, scale, None, dropout_p, training, 2 * salt)
    out_l = attention(k, q, vv, heads, scale, None, dropout_p, training, 2 * salt + 1)
    return out_v, out_l


# ---- input projections: GroupNorm on token-major activations, all feature levels into one token buffer -------------
# Off by default: the kernels are parity-checked on the GPU (tests/test_fused_gpu.py) and the model plumbing on CPU
# (tests/test_flat_levels_cpu.py), but the path has not had its A/B run inside the train step yet (in r01s4f a shape check
# kept it from engaging).  RLIPV2_GN_TOKENS=1 enables it.
_GN_TOKENS = os.environ.get("RLIPV2_GN_TOKENS", "0") != "0"


class _GroupNormTokensMulti(torch.autograd.Function):
    """[x_l: conv outputs [N, 256, H_l, W_l] in channels_last memory] -> GroupNorm(32) of each, written as rows of one
    [N, sum_l H_l W_l, 256] tensor (what `src.flatten(2).transpose(1, 2)` + `torch.cat` builds in the reference,
    dab_deformable/deformable_transformer.py:452-470) - csrc/fused_ops.cu gn_tok_* kernels."""

    @staticmethod
    def forward(ctx, eps, n_levels, *args):
        f = _fused()
        xs, gammas, betas = args[:n_levels], args[n_levels:2 * n_levels], args[2 * n_levels:]
        N, C = xs[0].shape[:2]
        hws = [x.shape[2] * x.shape[3] for x in xs]
        S = sum(hws)
        out = torch.empty((N, S, C), dtype=torch.float32, device=xs[0].device)
        saved, start = [], 0
        for x, g, b, hw in zip(xs, gammas, betas, hws):
            xt = x.permute(0, 2, 3, 1)                         # [N, H, W, C]: contiguous for channels_last tensors
            xt = (xt if xt.is_contiguous() else xt.contiguous()).view(N, hw, C)
            mean, rstd = f.groupnorm_tokens_fwd(xt, g, b, eps, out[:, start], S * C)
            saved += [xt, mean, rstd, g]
            start += hw
        ctx.save_for_backward(*saved)
        ctx.meta = (n_levels, N, C, S, [tuple(x.shape) for x in xs], hws)
        ctx.params = (gammas, betas)
        return out

    @staticmethod
    def backward(ctx, dy):
        f = _fused()
        n_levels, N, C, S, shapes, hws = ctx.meta
        saved = ctx.saved_tensors
        dy = dy if dy.is_contiguous() else dy.contiguous()
        gammas, betas = ctx.params
        dxs, dgs, dbs, start = [], [], [], 0
        for l in range(n_levels):
            xt, mean, rstd, g = saved[4 * l:4 * l + 4]
            gp, bp = gammas[l], betas[l]
            fuse = (getattr(gp, "_fuse_grad", False) and getattr(bp, "_fuse_grad", False) and gp.grad is not None
                    and bp.grad is not None and gp.grad.is_contiguous() and bp.grad.is_contiguous())
            dg = gp.grad if fuse else torch.zeros(C, dtype=torch.float32, device=dy.device)
            db = bp.grad if fuse else torch.zeros(C, dtype=torch.float32, device=dy.device)
            dx = f.groupnorm_tokens_bwd(dy[:, start], S * C, xt, mean, rstd, g, dg, db)
            _, _, H, W = shapes[l]
            dxs.append(dx.view(N, H, W, C).permute(0, 3, 1, 2))
            dgs.append(None if fuse else dg)
            dbs.append(None if fuse else db)
            start += hws[l]
        return (None, None) + tuple(dxs) + tuple(dgs) + tuple(dbs)


def group_norm_tokens_supported(features, convs, norms):
    """features: the tensors the projections' convolutions will read; convs / norms: the (Conv2d, GroupNorm) pairs"""
    return (_GN_TOKENS and _USE_FUSED
            and all(x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 for x in features)
            and all(isinstance(c, torch.nn.Conv2d) and c.out_channels == 256 for c in convs)
            and all(isinstance(n, torch.nn.GroupNorm) and n.num_groups == 32 and n.num_channels == 256 and n.affine
                    for n in norms))


def group_norm_tokens(xs, norms):
    """GroupNorm(32, 256) of every level's projected feature map, flattened to tokens and concatenated: [N, S, 256]"""
    eps = norms[0].eps
    assert all(n.eps == eps for n in norms)
    return _GroupNormTokensMulti.apply(eps, len(xs), *xs, *[n.weight for n in norms], *[n.bias for n in norms])


# ---- DAB decoder small-op chains (csrc/fused_ops.cu: box_refine_kernel, sine_embed_kernel) ---------------------
_SMALL_OPS = os.environ.get("RLIPV2_SMALL_OPS", "1") != "0"


def _inverse_sigmoid_torch(x, eps=1e-5):
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


class _BoxRefine(torch.autograd.Function):
    """sigmoid(delta + inverse_sigmoid(ref)) in one kernel; the gradient of `ref` (needed for the first decoder
    layer only, whose anchors are the learned `refpoint_embed`) is formed by torch from the same formula."""

    @staticmethod
    def forward(ctx, delta, ref, eps):
        d = delta if delta.is_contiguous() else delta.contiguous()
        r = ref if ref.is_contiguous() else ref.contiguous()
        y = _fused().box_refine(d, r, eps)
        ctx.eps = eps
        ctx.save_for_backward(y, r)
        return y

    @staticmethod
    def backward(ctx, g):
        y, r = ctx.saved_tensors
        gz = torch.ops.aten.sigmoid_backward(g.contiguous(), y)
        gref = None
        if ctx.needs_input_grad[1]:
            with torch.enable_grad():
                rr = r.detach().requires_grad_(True)
                (gref,) = torch.autograd.grad(_inverse_sigmoid_torch(rr, ctx.eps), rr, gz)
        return gz, gref, None


def box_refine(delta, ref, eps=1e-5):
    """(delta + inverse_sigmoid(ref)).sigmoid() - the iterative box refinement of the DAB decoder
    (dab_deformable/deformable_transformer.py:1511-1541; inverse_sigmoid: util/misc.py:460-464)"""
    if _SMALL_OPS and _USE_FUSED and delta.is_cuda and delta.dtype == torch.float32 and delta.shape == ref.shape \
            and delta.numel() > 0:
        return _BoxRefine.apply(delta, ref, eps)
    return (delta + _inverse_sigmoid_torch(ref, eps)).sigmoid()


def _sine_embed_torch(pos_tensor):
    n = pos_tensor.size(-1)
    dim_t = torch.arange(128, dtype=torch.float32, device=pos_tensor.device)
    dim_t = 10000 ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / 128)
    # (x, y, ...) -> (y, x, ...) with slices only: an index list would become a host tensor + H2D copy, which a
    # CUDA-graph capture rejects
    yx = torch.cat((pos_tensor[..., 1:2], pos_tensor[..., 0:1], pos_tensor[..., 2:]), dim=-1)
    p = (yx * (2 * math.pi))[..., None] / dim_t                                   # [.., n, 128]
    emb = torch.stack((p[..., 0::2].sin(), p[..., 1::2].cos()), dim=-1)          # [.., n, 64, 2]
    return emb.flatten(-3)


class _SineEmbed(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos):
        n = pos.shape[-1]
        p2 = pos.reshape(-1, n)
        p2 = p2 if p2.is_contiguous() else p2.contiguous()
        ctx.save_for_backward(pos)
        return _fused().sine_embed(p2).view(*pos.shape[:-1], n * 128)

    @staticmethod
    def backward(ctx, g):
        (pos,) = ctx.saved_tensors
        with torch.enable_grad():
            pp = pos.detach().requires_grad_(True)
            (gp,) = torch.autograd.grad(_sine_embed_torch(pp), pp, g)
        return gp


def sine_embed(pos_tensor):
    """[..., 2|4] normalised (x, y[, w, h]) -> [..., 256|512] sine embedding in the order (y, x[, w, h]); 128
    features each, temperature 10000 (gen_sineembed_for_position, deformable_transformer.py:1777-1802)."""
    if pos_tensor.size(-1) not in (2, 4):
        raise ValueError("Unknown pos_tensor shape(-1):{}".format(pos_tensor.size(-1)))
    if _SMALL_OPS and _USE_FUSED and pos_tensor.is_cuda and pos_tensor.dtype == torch.float32 and pos_tensor.numel() > 0:
        return _SineEmbed.apply(pos_tensor)
    return _sine_embed_torch(pos_tensor)
