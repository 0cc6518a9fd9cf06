"""Dense-contraction seam of the hot path.

Every GEMM-shaped op of ALIF, the RobertaLayer stack and the deformable encoder/decoder FFNs goes
through these functions, so the module code stays independent of how the contraction is executed.
Round 1 routes them to cuBLAS/cuDNN through torch (fp32 or TF32, see `set_matmul_precision`);
the hand-written tcgen05 kernels replace the bodies without touching the callers.
"""
import torch
import torch.nn.functional as F

_PRECISION = "fp32"


def set_matmul_precision(mode: str):
    """'fp32' = IEEE fp32 SGEMM (parity tests); 'tf32' = TF32 tensor-core products with fp32
    accumulation (what torch 1.10, the reference's pinned version, did by default on Ampere+)."""
    global _PRECISION
    assert mode in ("fp32", "tf32")
    _PRECISION = mode
    torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
    torch.backends.cudnn.allow_tf32 = mode == "tf32"


def matmul_precision():
    return _PRECISION


def linear(x, weight, bias=None):
    return F.linear(x, weight, bias)


def linear_relu(x, weight, bias=None):
    return F.relu(F.linear(x, weight, bias))


def linear_gelu(x, weight, bias=None):
    return F.gelu(F.linear(x, weight, bias))


def layer_norm(x, weight, bias, eps):
    return F.layer_norm(x, (x.shape[-1],), weight, bias, eps)


def add_layer_norm(x, residual, weight, bias, eps):
    """LayerNorm(x + residual)."""
    return F.layer_norm(x + residual, (x.shape[-1],), weight, bias, eps)
