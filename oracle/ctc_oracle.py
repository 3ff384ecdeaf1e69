"""numpy float64 restatement of the reference CTC loss.  TEST INFRASTRUCTURE ONLY.

The reference's CTC arithmetic is a third-party dependency that is not under /root/reference:
``torch.nn.CTCLoss`` (ATen ``ctc_loss``), reached from NeMo/nemo/collections/asr/losses/ctc.py:58,77-79
with ``blank = num_classes`` (last index, :46), ``zero_infinity=True`` and ``reduction='none'`` followed by
the NeMo-side ``mean_batch`` / ``mean_volume`` reduction (:52-66).  ``torch`` is UNPINNED in the reference
(NeMo/requirements/requirements.txt lists bare ``torch``); the authoring container has torch 2.11.0.

Published algorithm restated here (Graves et al. 2006, as implemented by ATen LossCTC):
  extended label l' = [blank, l1, blank, l2, ..., blank]           (S = 2U+1 states)
  alpha_0(0) = y_0(blank), alpha_0(1) = y_0(l1)
  alpha_t(s) = y_t(l'_s) * (alpha_{t-1}(s) + alpha_{t-1}(s-1) + [l'_s != blank and l'_s != l'_{s-2}] alpha_{t-1}(s-2))
  nll = -log(alpha_{T-1}(S-1) + alpha_{T-1}(S-2))
  beta symmetric; gradient convention = ATen's ``ctc_loss_backward``:
      grad_log_probs[t,c] = exp(lp[t,c]) - exp( logsumexp_{s: l'_s = c}(alpha_t(s)+beta_t(s)) + nll - lp[t,c] )
  which is only a true gradient after composition with log_softmax backward (the exp(lp) term then sums out);
  rows t >= input_length are zero; an infeasible sample (nll = inf) with zero_infinity gives loss 0, grad 0.

Pinned by: tests/golden/ref_kat.npz (reference tests/collections/asr/k2/test_ctc.py:85-188 costs + logits-grads)
and by torch.nn.CTCLoss itself (installed here and on the GPU box) on seeded random cases.
"""
from __future__ import annotations

import numpy as np


def ctc_loss_and_grad(
    log_probs: np.ndarray,
    targets: np.ndarray,
    input_lengths: np.ndarray,
    target_lengths: np.ndarray,
    blank: int,
    zero_infinity: bool = True,
    want_grad: bool = True,
):
    """log_probs [B,T,Vp] (NeMo layout, batch-major).  Returns nll[B], grad_log_probs[B,T,Vp] for
    d(sum_b nll_b) in ATen's convention (see module docstring)."""
    lp = np.asarray(log_probs, dtype=np.float64)
    B, maxT, Vp = lp.shape
    nll = np.zeros(B)
    grad = np.zeros_like(lp) if want_grad else None
    for b in range(B):
        T = int(input_lengths[b])
        U = int(target_lengths[b])
        S = 2 * U + 1
        ext = np.full(S, blank, dtype=np.int64)
        ext[1::2] = np.asarray(targets[b, :U], dtype=np.int64)
        skip = np.zeros(S, dtype=bool)  # may take s-2
        skip[2:] = (ext[2:] != blank) & (ext[2:] != ext[:-2])
        alpha = np.full((T, S), -np.inf)
        alpha[0, 0] = lp[b, 0, blank]
        if S > 1:
            alpha[0, 1] = lp[b, 0, ext[1]]
        for t in range(1, T):
            a0 = alpha[t - 1]
            a1 = np.concatenate(([-np.inf], a0[:-1]))
            a2 = np.concatenate(([-np.inf, -np.inf], a0[:-2]))[:S]
            a2 = np.where(skip, a2, -np.inf)
            alpha[t] = np.logaddexp(np.logaddexp(a0, a1), a2) + lp[b, t, ext]
        ll = alpha[T - 1, S - 1]
        if S > 1:
            ll = np.logaddexp(ll, alpha[T - 1, S - 2])
        nll_b = -ll
        if np.isinf(nll_b):
            nll[b] = 0.0 if zero_infinity else np.inf
            if want_grad and not zero_infinity:
                grad[b] = np.nan
            continue
        nll[b] = nll_b
        if not want_grad:
            continue
        beta = np.full((T, S), -np.inf)
        beta[T - 1, S - 1] = lp[b, T - 1, blank]
        if S > 1:
            beta[T - 1, S - 2] = lp[b, T - 1, ext[S - 2]]
        skip_fwd = np.zeros(S, dtype=bool)  # may go to s+2
        skip_fwd[:-2] = skip[2:]
        for t in range(T - 2, -1, -1):
            b0 = beta[t + 1]
            b1 = np.concatenate((b0[1:], [-np.inf]))
            b2 = np.concatenate((b0[2:], [-np.inf, -np.inf]))[-S:] if S > 1 else np.full(1, -np.inf)
            b2 = np.where(skip_fwd, b2, -np.inf)
            beta[t] = np.logaddexp(np.logaddexp(b0, b1), b2) + lp[b, t, ext]
        ab = alpha + beta  # [T,S]
        lcab = np.full((T, Vp), -np.inf)
        for s in range(S):
            lcab[:, ext[s]] = np.logaddexp(lcab[:, ext[s]], ab[:, s])
        grad[b, :T, :] = np.exp(lp[b, :T, :]) - np.exp(lcab + nll_b - lp[b, :T, :])
    return nll, grad


def reduce_losses(losses: np.ndarray, target_lengths: np.ndarray, reduction: str):
    """CTCLoss.reduce + torch reductions, losses/ctc.py:45-66."""
    if reduction == "mean_batch":
        return losses.mean()
    if reduction == "mean_volume":
        return losses.sum() / target_lengths.sum()
    if reduction == "sum":
        return losses.sum()
    if reduction == "mean":  # torch: divide by clamp_min(target_len, 1) then mean
        return (losses / np.maximum(target_lengths, 1)).mean()
    return losses
