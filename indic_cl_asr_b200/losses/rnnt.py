"""Transducer loss with the reference's interfaces, computed by hand-written sm_100a kernels.

Mirrors (paths relative to /root/reference/NeMo/nemo/collections/asr/):
  RNNTLossNumba       parts/numba/rnnt_loss/rnnt_pytorch.py:390-437  (blank, reduction, fastemit_lambda, clamp)
  _RNNTNumba          parts/numba/rnnt_loss/rnnt_pytorch.py:40-91
  certify_inputs      parts/numba/rnnt_loss/rnnt_pytorch.py:584-632  (same exceptions, same messages)
  RNNTLoss            losses/rnnt.py:333-508                          (facade: casts, narrowing, reduce)
  resolve_rnnt_loss   losses/rnnt.py:206-330                          ('default' -> warprnnt_numba, :158)

Differences that are deliberate:
  * CUDA only; a CPU tensor raises (the reference's CPU path is the test oracle, not the product).
  * the gradient is produced in backward() by one kernel that folds grad_output in, instead of being
    materialised in forward() and rescaled in place (rnnt_pytorch.py:58,87-91): one [B,T,U,V] write, no
    zero-fill, and double-backward-free just like the reference.
  * no stream synchronisation inside the loss (reference: gpu_rnnt.py:229).
"""
from __future__ import annotations

from typing import List, Optional

import torch

from .. import _lib
from .._typecheck import kwargs_only

__all__ = ["RNNTLoss", "RNNTLossNumba", "rnnt_loss", "certify_inputs", "resolve_rnnt_loss"]


def _check_type(var, t, name):
    if var.dtype is not t:
        raise TypeError("{} must be {}".format(name, t))


def _check_contiguous(var, name):
    if not var.is_contiguous():
        raise ValueError("{} must be contiguous".format(name))


def _check_dim(var, dim, name):
    if len(var.shape) != dim:
        raise ValueError("{} must be {}D".format(name, dim))


def certify_inputs(log_probs, labels, lengths, label_lengths, check_lengths: bool = True):
    """Same checks, order and messages as the reference (rnnt_pytorch.py:599-632).

    ``check_lengths=False`` skips only the two max() comparisons, which need a device->host sync."""
    _check_type(labels, torch.int64, "labels")
    _check_type(label_lengths, torch.int64, "label_lengths")
    _check_type(lengths, torch.int64, "lengths")
    _check_contiguous(log_probs, "log_probs")
    _check_contiguous(labels, "labels")
    _check_contiguous(label_lengths, "label_lengths")
    _check_contiguous(lengths, "lengths")

    if lengths.shape[0] != log_probs.shape[0]:
        raise ValueError(
            f"Must have a length per example. "
            f"Given lengths dim: {lengths.shape[0]}, "
            f"Log probs dim : {log_probs.shape[0]}"
        )
    if label_lengths.shape[0] != log_probs.shape[0]:
        raise ValueError(
            "Must have a label length per example. "
            f"Given label lengths dim : {label_lengths.shape[0]}, "
            f"Log probs dim : {log_probs.shape[0]}"
        )

    _check_dim(log_probs, 4, "log_probs")
    _check_dim(labels, 2, "labels")
    _check_dim(lengths, 1, "lenghts")
    _check_dim(label_lengths, 1, "label_lenghts")
    if check_lengths:
        max_T = torch.max(lengths)
        max_U = torch.max(label_lengths)
        T, U = log_probs.shape[1:3]
        if T != max_T:
            raise ValueError(f"Input length mismatch! Given T: {T}, Expected max T from input lengths: {max_T}")
        if U != max_U + 1:
            raise ValueError(f"Output length mismatch! Given U: {U}, Expected max U from target lengths: {max_U} + 1")


class _RNNTLossB200(torch.autograd.Function):
    """Per-sample costs on raw logits.  forward = LSE+gather kernel and alpha/beta wavefront;
    backward = softmax-fused gradient kernel (reference _RNNTNumba, rnnt_pytorch.py:40-91)."""

    @staticmethod
    def forward(ctx, acts, labels, act_lens, label_lens, blank, fastemit_lambda, clamp, check_lengths):
        _lib.require_cuda(acts, "acts")
        certify_inputs(acts, labels, act_lens, label_lens, check_lengths)
        if clamp < 0:
            raise ValueError("`clamp` must be 0.0 or positive float value.")
        B, T, U1, Vp = acts.shape
        L = _lib.lib()
        ws_bytes = L.clasr_rnnt_workspace_bytes(B, T, U1)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=acts.device)
        costs = torch.empty(B, dtype=torch.float32, device=acts.device)
        with torch.cuda.device(acts.device):
            st = L.clasr_rnnt_loss_fwd(
                acts.data_ptr(), labels.data_ptr(), act_lens.data_ptr(), label_lens.data_ptr(), B, T, U1, Vp,
                int(blank), float(fastemit_lambda), costs.data_ptr(), ws.data_ptr(), ws_bytes,
                _lib.stream_ptr(acts.device))
        _lib.check(st, "rnnt_loss_fwd")
        ctx.save_for_backward(acts, labels, act_lens, label_lens, ws)
        ctx.args = (int(blank), float(fastemit_lambda), float(clamp), ws_bytes)
        return costs

    @staticmethod
    def backward(ctx, grad_output):
        acts, labels, act_lens, label_lens, ws = ctx.saved_tensors
        blank, fastemit_lambda, clamp, ws_bytes = ctx.args
        B, T, U1, Vp = acts.shape
        go = grad_output.contiguous().to(torch.float32).view(-1)
        grads = torch.empty_like(acts)
        with torch.cuda.device(acts.device):
            st = _lib.lib().clasr_rnnt_loss_bwd(
                acts.data_ptr(), labels.data_ptr(), act_lens.data_ptr(), label_lens.data_ptr(), B, T, U1, Vp, blank,
                fastemit_lambda, clamp, go.data_ptr(), grads.data_ptr(), ws.data_ptr(), ws_bytes,
                _lib.stream_ptr(acts.device))
        _lib.check(st, "rnnt_loss_bwd")
        return grads, None, None, None, None, None, None, None


def rnnt_loss(acts, labels, act_lens, label_lens, blank=0, reduction="mean", fastemit_lambda: float = 0.0,
              clamp: float = 0.0, check_lengths: bool = True):
    """Functional form (reference rnnt_pytorch.py:345-387)."""
    if acts.dtype != torch.float32:
        acts = acts.float()
    costs = _RNNTLossB200.apply(acts, labels, act_lens, label_lens, blank, fastemit_lambda, clamp, check_lengths)
    if reduction in ["sum", "mean"]:
        costs = costs.sum().unsqueeze(-1)
        if reduction == "mean":
            costs = costs / acts.size(0)
    return costs


class RNNTLossNumba(torch.nn.Module):
    """Drop-in for the reference's RNNTLossNumba (rnnt_pytorch.py:390-437); the name is kept so that
    ``resolve_rnnt_loss('warprnnt_numba')`` call sites keep working."""

    def __init__(self, blank=0, reduction="mean", fastemit_lambda: float = 0.0, clamp: float = -1,
                 check_lengths: bool = True):
        super().__init__()
        self.blank = blank
        self.fastemit_lambda = fastemit_lambda
        self.clamp = float(clamp) if clamp > 0 else 0.0
        self.reduction = reduction
        self.check_lengths = check_lengths

    def forward(self, acts, labels, act_lens, label_lens):
        return rnnt_loss(acts, labels, act_lens, label_lens, self.blank, self.reduction, self.fastemit_lambda,
                         self.clamp, self.check_lengths)


RNNTLossB200 = RNNTLossNumba

# Names the reference's resolver knows that map onto the plain transducer loss (losses/rnnt.py:96-158,243-247).
_PLAIN_RNNT = ("default", "warprnnt_numba", "b200")
_OTHER_REFERENCE_LOSSES = ("warprnnt", "pytorch", "multiblank_rnnt", "multiblank_rnnt_pytorch", "graph_rnnt",
                           "graph_w_transducer", "tdt", "tdt_pytorch")


def resolve_rnnt_loss(loss_name: str, blank_idx: int, loss_kwargs: dict = None) -> torch.nn.Module:
    if loss_name not in _PLAIN_RNNT + _OTHER_REFERENCE_LOSSES:
        raise ValueError(f"Provided `loss_name` {loss_name} not in list of available RNNT losses \n"
                         f"{_PLAIN_RNNT + _OTHER_REFERENCE_LOSSES}")
    if loss_name in _OTHER_REFERENCE_LOSSES:
        raise NotImplementedError(
            f"loss_name={loss_name!r} is outside the hot path this library accelerates (plain RNNT, the reference's "
            "default 'warprnnt_numba'); multi-blank / TDT / k2 graph losses are out of scope (SURVEY.md §2.4 K8/K9)")
    loss_kwargs = {} if loss_kwargs is None else dict(loss_kwargs)
    fastemit_lambda = loss_kwargs.pop("fastemit_lambda", 0.0)
    clamp = loss_kwargs.pop("clamp", -1.0)
    check_lengths = loss_kwargs.pop("check_lengths", True)
    if loss_kwargs:
        raise ValueError(f"Loss function `{loss_name}` was provided with unused kwargs: {sorted(loss_kwargs)}")
    return RNNTLossNumba(blank=blank_idx, reduction="none", fastemit_lambda=fastemit_lambda, clamp=clamp,
                         check_lengths=check_lengths)


class RNNTLoss(torch.nn.Module):
    """Facade with the reference's constructor, attributes and call contract (losses/rnnt.py:333-508)."""

    def __init__(self, num_classes, reduction: str = "mean_batch", loss_name: str = "default", loss_kwargs=None):
        super().__init__()
        if reduction not in [None, "mean", "sum", "mean_batch", "mean_volume"]:
            raise ValueError("`reduction` must be one of [mean, sum, mean_batch, mean_volume]")
        self._blank = num_classes
        self.reduction = reduction
        self._loss = resolve_rnnt_loss(loss_name, blank_idx=self._blank, loss_kwargs=loss_kwargs)
        self._force_float32 = True

    @property
    def fastemit_lambda(self) -> float:
        return self._loss.fastemit_lambda

    @property
    def clamp(self) -> float:
        return self._loss.clamp

    def reduce(self, losses, target_lengths):
        if isinstance(losses, List):
            losses = torch.cat(losses, 0)
            target_lengths = torch.cat(target_lengths, 0)
        if self.reduction == "mean_batch":
            losses = losses.mean()
        elif self.reduction == "mean":
            losses = torch.div(losses, target_lengths).mean()
        elif self.reduction == "sum":
            losses = losses.sum()
        elif self.reduction == "mean_volume":
            losses = losses.sum() / target_lengths.sum()
        return losses

    @kwargs_only
    def forward(self, log_probs, targets, input_lengths, target_lengths):
        targets = targets.long()
        input_lengths = input_lengths.long()
        target_lengths = target_lengths.long()

        max_logit_len = input_lengths.max()
        max_targets_len = target_lengths.max()

        if log_probs.dtype != torch.float32:
            log_probs = log_probs.float()

        # narrow to the true maxima (losses/rnnt.py:475-484)
        if log_probs.shape[1] != max_logit_len:
            log_probs = log_probs.narrow(dim=1, start=0, length=int(max_logit_len)).contiguous()
        if not targets.is_contiguous():
            targets = targets.contiguous()
        if targets.shape[1] != max_targets_len:
            targets = targets.narrow(dim=1, start=0, length=int(max_targets_len)).contiguous()

        loss_reduction = self._loss.reduction
        self._loss.reduction = None
        loss = self._loss(acts=log_probs, labels=targets, act_lens=input_lengths, label_lens=target_lengths)
        self._loss.reduction = loss_reduction

        if self.reduction is not None:
            loss = self.reduce(loss, target_lengths)
        return loss
