"""Linear layer y = x W^T + b on the tcgen05 GEMM (C ABI: clasr_linear_fwd / clasr_linear_bwd).

Used for the joint's enc / pred projections (reference NeMo modules/rnnt.py:1563-1585; layers built :1679-1680) and
for the kernel-size-1 Conv1d of the CTC head (modules/conv_asr.py:444-446, 467-469).  Operands are split into bf16
hi + lo and multiplied as hi.hi + hi.lo + lo.hi with fp32 accumulation in tensor memory ("bf16x3": ~2^-17 relative
per product, fp32-grade) — with TF32 disabled, torch runs the same layers as SIMT sgemm / cuDNN convolutions.
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["linear_x3"]


def _is_transposed_view(x: torch.Tensor) -> bool:
    """True for ``base.transpose(1, 2)`` of a contiguous fp32 ``base`` [B, K, T] — how the joint and the CTC head receive the
    NeMo-layout encoder / prediction-network outputs (reference modules/rnnt.py:1457-1459, conv_asr.py:467)."""
    return (x.dim() == 3 and x.dtype == torch.float32 and x.shape[1] > 1 and x.shape[2] > 1 and x.stride(1) == 1
            and x.stride(2) == x.shape[1] and x.stride(0) == x.shape[1] * x.shape[2])


def _transpose_last2(src: torch.Tensor, B: int, R: int, C: int) -> torch.Tensor:
    """``src`` = contiguous fp32 [B, R, C] storage; returns a new contiguous [B, C, R] (clasr_transpose_last2)."""
    out = torch.empty(B, C, R, dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        _lib.check(_lib.lib().clasr_transpose_last2(src.data_ptr(), out.data_ptr(), B, R, C, _lib.stream_ptr(src.device)),
                   "transpose_last2")
    return out


def _rows(x: torch.Tensor, K: int):
    """The [M, K] row-major operand of ``x`` [..., K] and whether ``x`` was a transposed NeMo-layout view (then the input
    gradient goes back the same way instead of through ATen's generic strided copies: ~20 us each at B32/T250/D512)."""
    if _is_transposed_view(x):
        B, T, _ = x.shape
        return _transpose_last2(x, B, K, T).view(B * T, K), True     # storage of x is [B, K, T]
    return x.reshape(-1, K).contiguous().float(), False


def _input_grad(dx: torch.Tensor, xshape, transposed: bool) -> torch.Tensor:
    """``dx`` [M, K] back in the layout of the forward input: for a transposed view, a transposed view of a contiguous
    [B, K, T] buffer (autograd's transpose backward then hands the leaf a contiguous gradient)."""
    if not transposed:
        return dx.view(xshape)
    B, T, K = xshape
    return _transpose_last2(dx, B, T, K).transpose(1, 2)


class _LinearX3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, precision):
        _lib.require_cuda(x, "x")
        lead = x.shape[:-1]
        K = x.shape[-1]
        N = weight.shape[0]
        if weight.shape[1] != K:
            raise ValueError(f"linear: weight {tuple(weight.shape)} does not match input features {K}")
        x2, transposed = _rows(x, K)
        w = weight.contiguous().float()
        b = None if bias is None else bias.contiguous().float()
        M = x2.shape[0]
        prec = _lib.PREC[precision]
        L = _lib.lib()
        nbytes = L.clasr_linear_workspace_bytes(M, N, K, prec)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        y = torch.empty(M, N, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(L.clasr_linear_fwd(x2.data_ptr(), w.data_ptr(), _lib.ptr(b), y.data_ptr(), M, N, K, prec,
                                          ws.data_ptr(), nbytes, _lib.stream_ptr(x.device)), "linear_fwd")
        ctx.save_for_backward(ws)
        ctx.dims = (M, N, K, prec, nbytes, tuple(x.shape), bias is not None, transposed)
        return y.view(*lead, N)

    @staticmethod
    def backward(ctx, dy):
        (ws,) = ctx.saved_tensors
        M, N, K, prec, nbytes, xshape, has_bias, transposed = ctx.dims
        dy2 = dy.reshape(M, N).contiguous().float()
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], has_bias and ctx.needs_input_grad[2]
        dx = torch.empty(M, K, dtype=torch.float32, device=dy.device) if need_x else None
        dw = torch.empty(N, K, dtype=torch.float32, device=dy.device) if need_w else None
        db = torch.empty(N, dtype=torch.float32, device=dy.device) if need_b else None
        L = _lib.lib()
        with torch.cuda.device(dy.device):
            _lib.check(L.clasr_linear_bwd(dy2.data_ptr(), _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(db), M, N, K, prec,
                                          ws.data_ptr(), nbytes, _lib.stream_ptr(dy.device)), "linear_bwd")
        return (_input_grad(dx, xshape, transposed) if need_x else None), dw, db, None


class _LinearF16Fwd(torch.autograd.Function):
    """Forward with fp16 halves (2^-22 per product: the pre-activation of a ReLU joint must be fp32-grade, see
    RNNTJoint._project), backward with the range-safe bf16 split: upstream gradients are not pre-scaled and would
    underflow fp16, and nothing downstream of the backward GEMMs has a kink.  x and W are re-split for the backward
    pass (clasr_gemm_ex splits its fp32 operands itself; a tcgen05 MMA cannot mix an fp16 with a bf16 operand)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        _lib.require_cuda(x, "x")
        lead, K, N = x.shape[:-1], x.shape[-1], weight.shape[0]
        if weight.shape[1] != K:
            raise ValueError(f"linear: weight {tuple(weight.shape)} does not match input features {K}")
        x2, transposed = _rows(x, K)
        w = weight.contiguous().float()
        b = None if bias is None else bias.contiguous().float()
        M = x2.shape[0]
        prec = _lib.PREC["fp16x3"]
        L = _lib.lib()
        nbytes = L.clasr_linear_workspace_bytes(M, N, K, prec)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        y = torch.empty(M, N, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(L.clasr_linear_fwd(x2.data_ptr(), w.data_ptr(), _lib.ptr(b), y.data_ptr(), M, N, K, prec,
                                          ws.data_ptr(), nbytes, _lib.stream_ptr(x.device)), "linear_fwd")
        ctx.save_for_backward(x2, w)
        ctx.dims = (M, N, K, tuple(x.shape), bias is not None, transposed)
        return y.view(*lead, N)

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        M, N, K, xshape, has_bias, transposed = ctx.dims
        dy2 = dy.reshape(M, N).contiguous().float()
        L = _lib.lib()
        prec = _lib.PREC["bf16x3"]
        dev = dy.device

        def gemm(A, B, m, n, k, a_t, b_t, splits):
            nbytes = L.clasr_gemm_workspace_bytes(m, n, k, prec)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            C = torch.empty(m, n, dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                _lib.check(L.clasr_gemm_ex(A.data_ptr(), B.data_ptr(), C.data_ptr(), m, n, k, a_t, b_t, splits, prec,
                                           ws.data_ptr(), nbytes, _lib.stream_ptr(dev)), "gemm_ex")
            return C

        dx = _input_grad(gemm(dy2, w, M, K, N, 0, 1, 1), xshape, transposed) if ctx.needs_input_grad[0] else None   # dy . W
        dw = gemm(dy2, x2, N, K, M, 1, 1, max(1, min(16, M // 1024))) if ctx.needs_input_grad[1] else None   # dy^T . x
        db = dy2.sum(0) if (has_bias and ctx.needs_input_grad[2]) else None
        return dx, dw, db


def linear_x3(x: torch.Tensor, weight: torch.Tensor, bias=None, precision: str = "bf16x3") -> torch.Tensor:
    """``torch.nn.functional.linear(x, weight, bias)`` on the tcgen05 tensor cores (CUDA tensors only).
    ``precision="fp16x3"``: fp16 halves in the forward GEMM only (operands must sit in fp16's normal range), bf16 split
    in the backward GEMMs."""
    if precision == "fp16x3":
        return _LinearF16Fwd.apply(x, weight, bias)
    return _LinearX3.apply(x, weight, bias, precision)
