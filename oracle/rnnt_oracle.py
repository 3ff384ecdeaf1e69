"""numpy float64 restatement of the reference transducer (RNNT) loss.  TEST INFRASTRUCTURE ONLY.

Follows, function by function:
  log-softmax denominators      utils/cuda_utils/reduce.py:121-248 (K1/K2: max, then -max-log sum exp)
  forward variables (alpha)     utils/cpu_utils/cpu_rnnt.py:246-276 ; cuda gpu_rnnt_kernel.py:73-172
  backward variables (beta)     utils/cpu_utils/cpu_rnnt.py:278-318 ; cuda gpu_rnnt_kernel.py:175-269
  gradient w.r.t. LOGITS        utils/cuda_utils/gpu_rnnt_kernel.py:272-407 (softmax-fused; FastEmit; clamp)
  costs                         utils/rnnt_helper.py:106-116  (cost = -llForward * (1 + fastemit_lambda))
(all paths relative to /root/reference/NeMo/nemo/collections/asr/parts/numba/rnnt_loss/).

The CPU path of the reference returns d/d(log-probs) and lets autograd apply the log-softmax
backward (cpu_rnnt.py:320-343, rnnt_pytorch.py:411-437); composed, that equals the softmax-fused
logits gradient restated here (checked against the reference itself in tests/test_oracle_*.py).

Pinned by: tests/golden/ref_kat.npz (the reference's own known-answer vectors) and
tests/golden/ref_rnnt_*.npz (outputs of the reference's RNNTLossNumba run in the authoring container).
"""
from __future__ import annotations

import numpy as np


def _logaddexp(a, b):
    return np.logaddexp(a, b)


def log_softmax_denominator(logits: np.ndarray) -> np.ndarray:
    """denom[...] = -max - log(sum(exp(x - max)))  (negative log-sum-exp), reduce.py:186-248."""
    x = np.asarray(logits, dtype=np.float64)
    m = x.max(axis=-1)
    return -m - np.log(np.exp(x - m[..., None]).sum(axis=-1))


def alphas_betas(lp_blank: np.ndarray, lp_label: np.ndarray, T: int, U1: int):
    """Forward / backward variables over one utterance's T x U1 lattice.

    lp_blank[t,u] = log P(blank | t,u);  lp_label[t,u] = log P(label_u | t,u) (u < U1-1).
    Returns alpha[T,U1], beta[T,U1], ll_forward, ll_backward.
    """
    alpha = np.full((T, U1), -np.inf)
    beta = np.full((T, U1), -np.inf)
    alpha[0, 0] = 0.0
    for t in range(T):
        for u in range(U1):
            if u == 0 and t > 0:
                alpha[t, 0] = alpha[t - 1, 0] + lp_blank[t - 1, 0]
            if t == 0 and u > 0:
                alpha[0, u] = alpha[0, u - 1] + lp_label[0, u - 1]
            if t > 0 and u > 0:
                no_emit = alpha[t - 1, u] + lp_blank[t - 1, u]
                emit = alpha[t, u - 1] + lp_label[t, u - 1]
                alpha[t, u] = _logaddexp(emit, no_emit)
    ll_f = alpha[T - 1, U1 - 1] + lp_blank[T - 1, U1 - 1]

    beta[T - 1, U1 - 1] = lp_blank[T - 1, U1 - 1]
    for t in range(T - 1, -1, -1):
        for u in range(U1 - 1, -1, -1):
            if u == U1 - 1 and t < T - 1:
                beta[t, u] = beta[t + 1, u] + lp_blank[t, u]
            if t == T - 1 and u < U1 - 1:
                beta[t, u] = beta[t, u + 1] + lp_label[t, u]
            if t < T - 1 and u < U1 - 1:
                no_emit = beta[t + 1, u] + lp_blank[t, u]
                emit = beta[t, u + 1] + lp_label[t, u]
                beta[t, u] = _logaddexp(emit, no_emit)
    ll_b = beta[0, 0]
    return alpha, beta, ll_f, ll_b


def alphas_betas_diag(lp_blank: np.ndarray, lp_label: np.ndarray, T: int, U1: int):
    """Same recursion, vectorised over anti-diagonals (used for the larger oracle cases)."""
    alpha = np.full((T, U1), -np.inf)
    beta = np.full((T, U1), -np.inf)
    alpha[0, 0] = 0.0
    for n in range(1, T + U1 - 1):
        u = np.arange(max(0, n - (T - 1)), min(U1 - 1, n) + 1)
        t = n - u
        no_emit = np.full(u.shape, -np.inf)
        emit = np.full(u.shape, -np.inf)
        m = t > 0
        no_emit[m] = alpha[t[m] - 1, u[m]] + lp_blank[t[m] - 1, u[m]]
        m = u > 0
        emit[m] = alpha[t[m], u[m] - 1] + lp_label[t[m], u[m] - 1]
        alpha[t, u] = _logaddexp(emit, no_emit)
    ll_f = alpha[T - 1, U1 - 1] + lp_blank[T - 1, U1 - 1]
    beta[T - 1, U1 - 1] = lp_blank[T - 1, U1 - 1]
    for n in range(T + U1 - 3, -1, -1):
        u = np.arange(max(0, n - (T - 1)), min(U1 - 1, n) + 1)
        t = n - u
        no_emit = np.full(u.shape, -np.inf)
        emit = np.full(u.shape, -np.inf)
        m = t < T - 1
        no_emit[m] = beta[t[m] + 1, u[m]] + lp_blank[t[m], u[m]]
        m = u < U1 - 1
        emit[m] = beta[t[m], u[m] + 1] + lp_label[t[m], u[m]]
        beta[t, u] = _logaddexp(emit, no_emit)
    return alpha, beta, ll_f, beta[0, 0]


def rnnt_loss_and_grad(
    logits: np.ndarray,
    labels: np.ndarray,
    act_lens: np.ndarray,
    label_lens: np.ndarray,
    blank: int,
    fastemit_lambda: float = 0.0,
    clamp: float = 0.0,
    want_grad: bool = True,
    return_lattice: bool = False,
):
    """costs[B] and d(sum costs)/d(logits) [B,T,U1,Vp] in float64.

    logits are RAW joint outputs (the CUDA path of the reference consumes raw logits,
    rnnt_pytorch.py:411-437); padded cells (t >= T_b or u > U_b) get exactly-zero gradient
    (gpu_rnnt_kernel.py:343).
    """
    x = np.asarray(logits, dtype=np.float64)
    B, maxT, maxU1, Vp = x.shape
    labels = np.asarray(labels)
    denom = log_softmax_denominator(x)  # [B,T,U1]
    costs = np.zeros(B)
    grads = np.zeros_like(x) if want_grad else None
    lattice = []
    for b in range(B):
        T = int(act_lens[b])
        U1 = int(label_lens[b]) + 1
        lab = labels[b, : U1 - 1].astype(np.int64)
        logp = x[b, :T, :U1, :] + denom[b, :T, :U1, None]  # log-probs [T,U1,Vp]
        lp_blank = logp[:, :, blank]
        lp_label = np.full((T, U1), -np.inf)
        if U1 > 1:
            lp_label[:, : U1 - 1] = np.take_along_axis(
                logp[:, : U1 - 1, :], lab[None, :, None].repeat(T, 0), axis=2
            )[:, :, 0]
        fn = alphas_betas if T * U1 <= 4096 else alphas_betas_diag
        alpha, beta, ll_f, ll_b = fn(lp_blank, lp_label, T, U1)
        costs[b] = -ll_f * (1.0 + fastemit_lambda)
        if return_lattice:
            lattice.append((alpha, beta, ll_f, ll_b))
        if not want_grad:
            continue
        # gpu_rnnt_kernel.py:351-396
        ab = alpha + beta  # [T,U1]
        g = np.exp(ab[:, :, None] + logp - ll_f)
        if fastemit_lambda > 0.0 and U1 > 1:
            fe = fastemit_lambda * np.exp(
                alpha[:, : U1 - 1, None] + lp_label[:, : U1 - 1, None] + beta[:, 1:, None] + logp[:, : U1 - 1, :] - ll_f
            )
            g[:, : U1 - 1, :] += fe
        # final blank
        g[T - 1, U1 - 1, blank] -= np.exp(alpha[T - 1, U1 - 1] + lp_blank[T - 1, U1 - 1] - ll_f)
        # blank transitions t < T-1
        if T > 1:
            g[: T - 1, :, blank] -= np.exp(alpha[: T - 1, :] + lp_blank[: T - 1, :] - ll_f + beta[1:, :])
        # label transitions u < U1-1
        if U1 > 1:
            val = np.exp(np.log1p(fastemit_lambda) + alpha[:, : U1 - 1] + lp_label[:, : U1 - 1] - ll_f + beta[:, 1:])
            tt, uu = np.meshgrid(np.arange(T), np.arange(U1 - 1), indexing="ij")
            np.subtract.at(g, (tt, uu, lab[None, :].repeat(T, 0)), val)
        if clamp > 0.0:
            g = np.clip(g, -clamp, clamp)
        grads[b, :T, :U1, :] = g
    if return_lattice:
        return costs, grads, lattice
    return costs, grads


def reduce_losses(losses: np.ndarray, target_lengths: np.ndarray, reduction):
    """RNNTLoss.reduce, losses/rnnt.py:422-437."""
    if reduction == "mean_batch":
        return losses.mean()
    if reduction == "mean":
        return (losses / target_lengths).mean()
    if reduction == "sum":
        return losses.sum()
    if reduction == "mean_volume":
        return losses.sum() / target_lengths.sum()
    return losses
