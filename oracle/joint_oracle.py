"""torch-CPU restatement of the reference RNNTJoint maths and fused sub-batch loop.  TEST INFRASTRUCTURE ONLY.

Follows NeMo/nemo/collections/asr/modules/rnnt.py:
  projections                 :1563-1585  (enc: Linear(D_enc,H), pred: Linear(D_pred,H))
  joint_after_projection      :1587-1665  (f.unsqueeze(2) + g.unsqueeze(1) -> act -> [dropout] -> Linear(H,V+1);
                                           log_softmax only on CPU tensors or when log_softmax=True)
  fused sub-batch loop        :1403-1561  (narrow each sub-batch to its own max T / max U+1, per-sample
                                           losses with reduction=None, then RNNTLoss.reduce over the batch)
and rnnt_abstract.py:70-100 (joint = joint_after_projection(project_encoder(f), project_prednet(g))).

The transducer loss inside is oracle/rnnt_oracle.py wrapped as an autograd Function (float64 lattice),
so parameter gradients come out of ordinary torch autograd on CPU.

Pinned by tests/golden/ref_joint_*.npz: outputs of the reference's own RNNTJoint + RNNTLoss(+RNNTLossNumba)
source executed in the authoring container (oracle/ref_import.py, oracle/gen_golden.py).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import rnnt_oracle

_ACTS = {"relu": torch.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid}


class _RNNTLossFn(torch.autograd.Function):
    """Per-sample transducer costs on raw logits; backward = softmax-fused logits gradient."""

    @staticmethod
    def forward(ctx, logits, labels, act_lens, label_lens, blank, fastemit_lambda, clamp):
        costs, grads = rnnt_oracle.rnnt_loss_and_grad(
            logits.detach().numpy(), labels.numpy(), act_lens.numpy(), label_lens.numpy(), blank,
            fastemit_lambda=fastemit_lambda, clamp=clamp, want_grad=True,
        )
        ctx.grads = torch.from_numpy(grads).to(logits.dtype)
        return torch.from_numpy(costs).to(logits.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        return ctx.grads * grad_out.view(-1, 1, 1, 1), None, None, None, None, None, None


def rnnt_loss(logits, labels, act_lens, label_lens, blank, fastemit_lambda=0.0, clamp=0.0):
    return _RNNTLossFn.apply(logits, labels.long(), act_lens.long(), label_lens.long(), blank, fastemit_lambda, clamp)


def joint_logits(enc_out, pred_out, p: Dict[str, torch.Tensor], activation: str, language_ids=None):
    """enc_out [B,T,D_enc], pred_out [B,U1,D_pred] (already transposed as in forward :1388-1391).

    ``language_ids`` (one key per utterance) selects the multilingual head (:1627-1639): ``p`` then holds
    ``out.<lang>.weight / .bias`` per language; a batch of one language takes that head for the whole batch, a mixed
    batch goes utterance by utterance — numerically the same thing, so one code path serves both."""
    f = torch.nn.functional.linear(enc_out, p["enc.weight"], p["enc.bias"])
    g = torch.nn.functional.linear(pred_out, p["pred.weight"], p["pred.bias"])
    inp = _ACTS[activation](f.unsqueeze(2) + g.unsqueeze(1))
    if language_ids is None:
        return torch.nn.functional.linear(inp, p["out.weight"], p["out.bias"])
    return torch.stack([torch.nn.functional.linear(x, p[f"out.{lang}.weight"], p[f"out.{lang}.bias"])
                        for x, lang in zip(inp, language_ids)])


def reduce_losses(losses, target_lengths, reduction):
    if reduction == "mean_batch":
        return losses.mean()
    if reduction == "mean":
        return torch.div(losses, target_lengths).mean()
    if reduction == "sum":
        return losses.sum()
    if reduction == "mean_volume":
        return losses.sum() / target_lengths.sum()
    return losses


def fused_joint_loss(
    encoder_outputs,  # [B, D_enc, T]  (NeMo layout)
    decoder_outputs,  # [B, D_pred, U1]
    encoder_lengths,
    transcripts,
    transcript_lengths,
    p: Dict[str, torch.Tensor],
    activation: str,
    num_classes: int,
    fused_batch_size: int,
    reduction: Optional[str] = "mean_batch",
    fastemit_lambda: float = 0.0,
    clamp: float = 0.0,
    return_sub_logits: bool = False,
    language_ids=None,
):
    enc = encoder_outputs.transpose(1, 2)
    dec = decoder_outputs.transpose(1, 2)
    B = enc.shape[0]
    losses, tls, subs = [], [], []
    for begin in range(0, B, fused_batch_size):
        end = min(begin + fused_batch_size, B)
        sl = slice(begin, end)
        el, tl = encoder_lengths[sl], transcript_lengths[sl]
        mt, mu = int(el.max()), int(tl.max())
        z = joint_logits(enc[sl, :mt], dec[sl, : mu + 1], p, activation,
                         None if language_ids is None else list(language_ids[begin:end]))
        if return_sub_logits:
            subs.append(z)
        # blank = num_classes (RNNTLoss._blank, losses/rnnt.py:416)
        losses.append(rnnt_loss(z, transcripts[sl, :mu].contiguous(), el, tl, num_classes, fastemit_lambda, clamp))
        tls.append(tl)
    losses = torch.cat(losses, 0)
    tls = torch.cat(tls, 0)
    out = reduce_losses(losses, tls, reduction)
    if return_sub_logits:
        return out, subs
    return out


def ctc_head(encoder_output, weight, bias, column_index=None):
    """ConvASRDecoder.forward, conv_asr.py:458-490: Conv1d(k=1) -> optional per-language column
    select (masked_select with a row-constant mask == index_select of the True columns) -> log_softmax.
    encoder_output [B, D, T] -> (log_probs [B,T,Vp], logits [B,T,Vp])."""
    z = torch.nn.functional.conv1d(encoder_output, weight, bias).transpose(1, 2)
    if column_index is not None:
        z = z.index_select(-1, column_index)
    return torch.nn.functional.log_softmax(z, dim=-1), z


def mas_objective(sub_logits, ctc_logits, mas_ctx: float):
    """cl_baseline_mas.py:258-265."""
    dec = (ctc_logits.flatten(end_dim=-2) ** 2).sum(dim=-1).mean()
    r = 0
    for z in sub_logits:
        r = r + (z.flatten(end_dim=-2) ** 2).sum(dim=-1).mean()
    r = r / len(sub_logits)
    return r * (1 - mas_ctx) + dec * mas_ctx
