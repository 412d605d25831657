"""Python face of the hand-written sm_100a kernels (``csrc/kernels``).

Every function here launches OUR kernels - there is no torch/cuBLAS fallback: if the
native module is missing on a GPU box the import fails loudly (the driver checks which
.so files were loaded).  Matrices handed to the TMA-fed GEMMs must be fp32 with unit inner
stride, a row pitch that is a multiple of 4 floats and a 16-byte aligned base
(``tma_ready``); tensors that are not get one padded staging copy (``as_tma``).
"""
from __future__ import annotations

import torch

try:
    from .. import _C  # built in-tree: python setup.py build_ext --inplace
except ImportError as e:  # pragma: no cover
    raise ImportError(
        "shallowspeed_b200._C (sm_100a kernels + runtime) is not built. Run "
        "`python setup.py build_ext --inplace` (or __graft_entry__.build())."
    ) from e


def _round_up(x, m):
    return (x + m - 1) // m * m


def tma_ready(t: torch.Tensor) -> bool:
    return (t.dim() == 2 and t.dtype == torch.float32 and t.is_cuda and (t.stride(1) == 1 or t.size(1) == 1)
            and t.stride(0) % 4 == 0 and t.stride(0) >= t.size(1) and t.data_ptr() % 16 == 0)


def empty_padded(rows: int, cols: int, device) -> torch.Tensor:
    """[rows, cols] view of a [rows, round_up(cols, 4)] buffer (TMA-legal row pitch)."""
    # zeros: padding columns may be read by 16-byte-granular TMA accesses and must stay finite
    return torch.zeros(rows, _round_up(cols, 4), dtype=torch.float32, device=device)[:, :cols]


def as_tma(t: torch.Tensor) -> torch.Tensor:
    if tma_ready(t):
        return t
    out = empty_padded(t.size(0), t.size(1), t.device)
    out.copy_(t)
    return out


def _bias_arg(bias, out_dims):
    """bias may be [1, out] / [out] with any element stride (arena bias column)."""
    if bias is None:
        return None, 0
    b = bias.reshape(-1) if bias.dim() == 1 else bias
    if b.dim() == 2:
        assert b.size(0) == 1 and b.size(1) == out_dims
        stride = b.stride(1)
    else:
        assert b.numel() == out_dims
        stride = b.stride(0)
    return b, (stride if out_dims > 1 else 1)


# ------------------------------------------------------------------ GEMMs
# precision: "tf32" = one tcgen05.mma.kind::tf32 pass (operands truncated to 10 mantissa bits);
#            "fp32" = 3xTF32: lo twins (x - trunc(x)) of both operands, three MMAs per k-slice, error ~2^-20.
PRECISION = "fp32"


def _split(precision):
    return (precision or PRECISION) in ("fp32", "fp32x3", "3xtf32")


def lo_twin(t: torch.Tensor) -> torch.Tensor:
    """lo = t - trunc_tf32(t), same padded layout as t (t must be tma_ready).  The split kernel writes the whole padded
    storage (rows x pitch), so the buffer needs no memset: one kernel per twin, not two."""
    base = torch.empty(t.size(0), t.stride(0), dtype=torch.float32, device=t.device)
    lo = base[:, : t.size(1)]
    assert lo.stride(0) == t.stride(0)
    _C.split_lo(t, lo)
    return lo


def linear_fwd(x, weight, bias=None, relu=False, out=None, precision=None, k_splits=0):
    """k_splits: 0 = plain kernel, -1 = planner decides, k >= 2 = force k k-splits (experimental wide-layer variant)."""
    x, weight = as_tma(x), as_tma(weight)
    rows, out_dims = x.size(0), weight.size(0)
    y = out if out is not None else empty_padded(rows, out_dims, x.device)
    b, bstride = _bias_arg(bias, out_dims)
    if _split(precision):
        _C.linear_fwd(x, weight, b, bstride, bool(relu), y, lo_twin(weight), lo_twin(x), None, int(k_splits))
    else:
        _C.linear_fwd(x, weight, b, bstride, bool(relu), y, None, None, None, int(k_splits))
    return y


def linear_dgrad(dz, weight, mask=None, out=None, precision=None, k_splits=0):
    dz, weight = as_tma(dz), as_tma(weight)
    dx = out if out is not None else empty_padded(dz.size(0), weight.size(1), dz.device)
    m = None if mask is None else as_tma(mask)
    if _split(precision):
        _C.linear_dgrad(dz, weight, m, dx, lo_twin(weight), lo_twin(dz), None, int(k_splits))
    else:
        _C.linear_dgrad(dz, weight, m, dx, None, None, None, int(k_splits))
    return dx


def linear_wgrad(dz, x, grad_w, accumulate=False, grad_b=None, weight=None, lr=0.0, fuse_sgd=False, precision=None):
    """grad_w[out, in] (+)= dz^T @ x ; grad_b (+)= colsum(dz).  With ``fuse_sgd`` the update goes straight into
    ``weight`` (TMA reduce-add of -lr * dW)."""
    dz, x = as_tma(dz), as_tma(x)
    b, bstride = _bias_arg(grad_b, dz.size(1))
    if _split(precision):
        _C.linear_wgrad(dz, x, grad_w, bool(accumulate), b, bstride, weight, float(lr), bool(fuse_sgd), lo_twin(dz), lo_twin(x))
    else:
        _C.linear_wgrad(dz, x, grad_w, bool(accumulate), b, bstride, weight, float(lr), bool(fuse_sgd), None, None)
    return grad_w


def linear_grad(grad_output, input, weight):
    """functional API: returns fresh (dX, dW, db)."""
    dx = linear_dgrad(grad_output, weight)
    out_dims, in_dims = weight.shape
    dW = empty_padded(out_dims, in_dims, weight.device)
    db = torch.empty(out_dims, dtype=torch.float32, device=weight.device)
    linear_wgrad(grad_output, input, dW, accumulate=False, grad_b=db)
    return dx, dW, db


def linear_bwd_accumulate(dout, x, w_block, g_block, in_dims):
    """Module path: dX = dout @ W ; G_block[:, :in] += dout^T x ; G_block[:, in] += colsum(dout)."""
    W = w_block[:, :in_dims]
    dx = linear_dgrad(dout, W)
    linear_wgrad(dout, x, g_block[:, :in_dims], accumulate=True, grad_b=g_block[:, in_dims])
    return dx


# ------------------------------------------------------------------ loss head / softmax
def softmax(logits):
    logits = as_tma(logits)
    p = empty_padded(logits.size(0), logits.size(1), logits.device)
    _C.loss_head(logits, None, p, None, None, 0.0)
    return p


def softmax_grad(grad_output, logits):
    logits, up = as_tma(logits), as_tma(grad_output)
    dz = empty_padded(logits.size(0), logits.size(1), logits.device)
    _C.softmax_grad(logits, up, dz)
    return dz


def loss_head_backward(logits, target, batch_size):
    logits, target = as_tma(logits), as_tma(target)
    r, c = logits.shape
    p, dz = empty_padded(r, c, logits.device), empty_padded(r, c, logits.device)
    loss = torch.zeros(1, dtype=torch.float32, device=logits.device)
    _C.loss_head(logits, target, p, dz, loss, 1.0 / batch_size)
    return dz, p, loss[0]


def mse_loss_grad(input, target, batch_size):
    x, t = input.contiguous(), target.contiguous()
    y = torch.empty_like(x)
    _C.axpby(x, t, y, 2.0 / batch_size, -2.0 / batch_size)
    return y


# ------------------------------------------------------------------ elementwise
def relu(x):
    xc = x.contiguous()
    y = torch.empty_like(xc)
    _C.relu_fwd(xc, y)
    return y


def relu_grad(grad_output, mask_or_output):
    """mask_or_output: bool mask, or any fp32 tensor whose sign encodes it (layer output)."""
    g = empty_padded(grad_output.size(0), grad_output.size(1), grad_output.device)
    g.copy_(grad_output)
    m = mask_or_output
    if m.dtype != torch.float32:
        m = m.to(torch.float32)
    _C.relu_mask_(g, as_tma(m))
    return g


def sgd_step_(weights_flat, grads_flat, lr):
    _C.sgd_(weights_flat, grads_flat, float(lr))


def count_correct(pred, target):
    correct = torch.zeros(1, dtype=torch.int32, device=pred.device)
    _C.argmax_correct(as_tma(pred), as_tma(target), correct)
    return correct


def gemm_kernel_count() -> int:
    return _C.gemm_kernel_count()
