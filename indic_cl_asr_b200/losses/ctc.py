"""CTC loss with the reference's interface, computed by hand-written sm_100a kernels.

Mirrors NeMo/nemo/collections/asr/losses/ctc.py:25-81 (``CTCLoss(nn.CTCLoss)``): ``blank = num_classes``
(last index), reductions ``none | mean | sum | mean_batch | mean_volume`` with the last two applied on the
NeMo side, keyword-only ``forward(log_probs[B,T,V+1], targets[B,U], input_lengths, target_lengths)``.
The reference transposes to [T,B,V+1] for ATen; the kernels here read the batch-major layout directly.
"""
from __future__ import annotations

import torch

from .. import _lib
from .._typecheck import kwargs_only

__all__ = ["CTCLoss", "ctc_loss"]


class _CTCLossB200(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_probs, targets, input_lengths, target_lengths, blank, zero_infinity):
        _lib.require_cuda(log_probs, "log_probs")
        if log_probs.dim() != 3:
            raise ValueError("log_probs must be 3D [B, T, D]")
        if log_probs.dtype != torch.float32:
            raise TypeError("log_probs must be torch.float32")
        # decided BEFORE .contiguous(): inside Function.forward grad mode is off, so the contiguous copy of a narrowed /
        # transposed input would report requires_grad = False and the beta pass would be skipped
        need_beta = 1 if ctx.needs_input_grad[0] else 0
        log_probs = log_probs.contiguous()
        targets = targets.contiguous()
        B, T, Vp = log_probs.shape
        if targets.dim() != 2 or targets.shape[0] != B:
            raise ValueError("targets must be 2D [B, U]")
        if input_lengths.shape[0] != B or target_lengths.shape[0] != B:
            raise ValueError("input_lengths and target_lengths must have one entry per batch element")
        maxU = int(targets.shape[1])
        L = _lib.lib()
        ws_bytes = L.clasr_ctc_workspace_bytes(B, T, maxU)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=log_probs.device)
        nll = torch.empty(B, dtype=torch.float32, device=log_probs.device)
        with torch.cuda.device(log_probs.device):
            st = L.clasr_ctc_loss_fwd(
                log_probs.data_ptr(), _lib.ptr(targets) if maxU > 0 else 0, int(targets.stride(0)) if maxU > 0 else 0,
                input_lengths.data_ptr(), target_lengths.data_ptr(), B, T, Vp, maxU, int(blank),
                int(bool(zero_infinity)), need_beta, nll.data_ptr(), ws.data_ptr(), ws_bytes,
                _lib.stream_ptr(log_probs.device))
        _lib.check(st, "ctc_loss_fwd")
        ctx.save_for_backward(log_probs, targets, input_lengths, target_lengths, ws)
        ctx.args = (int(blank), int(bool(zero_infinity)), ws_bytes, maxU, need_beta)
        return nll

    @staticmethod
    def backward(ctx, grad_output):
        log_probs, targets, input_lengths, target_lengths, ws = ctx.saved_tensors
        blank, zero_infinity, ws_bytes, maxU, had_beta = ctx.args
        if not had_beta:
            raise RuntimeError("CTC backward requested but forward ran without requires_grad")
        B, T, Vp = log_probs.shape
        go = grad_output.contiguous().to(torch.float32).view(-1)
        grad = torch.empty_like(log_probs)
        with torch.cuda.device(log_probs.device):
            st = _lib.lib().clasr_ctc_loss_bwd(
                log_probs.data_ptr(), _lib.ptr(targets) if maxU > 0 else 0, int(targets.stride(0)) if maxU > 0 else 0,
                input_lengths.data_ptr(), target_lengths.data_ptr(), B, T, Vp, maxU, blank, zero_infinity,
                go.data_ptr(), grad.data_ptr(), ws.data_ptr(), ws_bytes, _lib.stream_ptr(log_probs.device))
        _lib.check(st, "ctc_loss_bwd")
        return grad, None, None, None, None, None


def ctc_loss(log_probs, targets, input_lengths, target_lengths, blank: int, zero_infinity: bool = False):
    """Per-sample negative log-likelihoods [B] for batch-major log_probs [B,T,V+1]."""
    return _CTCLossB200.apply(log_probs, targets.long(), input_lengths.long(), target_lengths.long(), blank,
                              zero_infinity)


class CTCLoss(torch.nn.Module):
    def __init__(self, num_classes, zero_infinity=False, reduction="mean_batch"):
        super().__init__()
        self._blank = num_classes
        if reduction not in ["none", "mean", "sum", "mean_batch", "mean_volume"]:
            raise ValueError("`reduction` must be one of [mean, sum, mean_batch, mean_volume]")
        self.config_reduction = reduction
        if reduction == "mean_batch" or reduction == "mean_volume":
            self.reduction = "none"
            self._apply_reduction = True
        else:
            self.reduction = reduction
            self._apply_reduction = False
        self.blank = self._blank
        self.zero_infinity = zero_infinity

    def reduce(self, losses, target_lengths):
        if self.config_reduction == "mean_batch":
            losses = losses.mean()
        elif self.config_reduction == "mean_volume":
            losses = losses.sum() / target_lengths.sum()
        return losses

    @kwargs_only
    def forward(self, log_probs, targets, input_lengths, target_lengths):
        input_lengths = input_lengths.long()
        target_lengths = target_lengths.long()
        targets = targets.long()
        loss = ctc_loss(log_probs, targets, input_lengths, target_lengths, self._blank, self.zero_infinity)
        # torch.nn.CTCLoss reductions (the reference's non-NeMo reductions, losses/ctc.py:55-57)
        if self.reduction == "sum":
            loss = loss.sum()
        elif self.reduction == "mean":
            loss = (loss / target_lengths.clamp_min(1).to(loss.dtype)).mean()
        if self._apply_reduction:
            loss = self.reduce(loss, target_lengths)
        return loss
