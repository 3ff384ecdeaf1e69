"""Autograd binding of the fused joint + transducer loss (C ABI: clasr_joint_rnnt_fwd / _bwd).

Inputs are the joint's projections f = enc(encoder_outputs) [B,T,H] and g = pred(decoder_outputs) [B,U+1,H]
(reference modules/rnnt.py:1563-1585) and the output layer's weight/bias; the result is the per-sample
transducer cost vector that the reference obtains from joint_after_projection (:1587-1665) followed by
RNNTLoss with reduction=None (:1475-1508) — without the [B,T,U+1,V+1] logits ever existing in HBM.
"""
from __future__ import annotations

import os

import torch

from . import _lib

__all__ = ["fused_joint_rnnt_loss", "fused_joint_forward_stats", "fused_joint_sumsq", "LazySubLogits",
           "dropout_mask_reference"]


def _stash_limit_bytes(stash_gib=None) -> int:
    """Largest logits stash (bytes) a differentiated forward call may allocate.

    Default 0: the logits are NEVER written to HBM and the backward pass recomputes them tile-wise on the tensor cores
    (BASELINE.json north_star (1)).  ``stash_gib`` > 0 (``RNNTJoint(backward_mode="stash")``) opts into the faster
    mode that keeps the valid cells' logits and hidden activations between forward and backward (about 15 % less step
    time at B32/T250/U100/V1024 for 4 (V+1+H) bytes per lattice cell).  The environment variable CLASR_JOINT_STASH
    (GiB; "0" = recompute) overrides both — it exists for A/B measurements."""
    v = os.environ.get("CLASR_JOINT_STASH", "")
    gib = float(v) if v != "" else (0.0 if stash_gib is None else float(stash_gib))
    return int(gib * (1 << 30))


def _stash(f, B, T, U1, H, Vp, prec, needs_grad, stash_gib=None):
    """(tensor | None, nbytes): the logits / hidden-activation stash a differentiated forward call leaves for its
    backward call (include/clasr_b200.h: clasr_joint_stash_bytes)."""
    if not needs_grad:
        return None, 0
    limit = _stash_limit_bytes(stash_gib)
    if limit <= 0:
        return None, 0
    nbytes = _lib.lib().clasr_joint_stash_bytes(B, T, U1, H, Vp, prec)
    if nbytes == 0 or nbytes > limit:
        return None, 0
    return torch.empty(nbytes, dtype=torch.uint8, device=f.device), nbytes


def _ws(f, B, T, U1, H, Vp, prec):
    L = _lib.lib()
    nbytes = L.clasr_joint_workspace_bytes(B, T, U1, H, Vp, prec)
    return torch.empty(nbytes, dtype=torch.uint8, device=f.device), nbytes


class _FusedJointRNNT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f, g, weight, bias, labels, act_lens, label_lens, blank, activation, precision, fastemit_lambda,
                clamp, want_sumsq, dropout_p=0.0, dropout_seed=0, stash_gib=None):
        _lib.require_cuda(f, "f")
        if clamp < 0:
            raise ValueError("`clamp` must be 0.0 or positive float value.")
        needs_grad = any(ctx.needs_input_grad[:4])
        f = f.contiguous().float()
        g = g.contiguous().float()
        weight = weight.contiguous().float()
        bias = bias.contiguous().float()
        labels = labels.contiguous().long()
        act_lens = act_lens.contiguous().long()
        label_lens = label_lens.contiguous().long()
        B, T, H = f.shape
        U1 = g.shape[1]
        Vp = weight.shape[0]
        if g.shape[0] != B or g.shape[2] != H or weight.shape[1] != H or bias.shape[0] != Vp:
            raise ValueError("fused joint: inconsistent shapes")
        if labels.shape[0] != B or (U1 > 1 and labels.shape[1] < U1 - 1):
            raise ValueError("fused joint: transcripts must be [B, >= U]")
        if labels.shape[1] != U1 - 1:
            labels = labels[:, : U1 - 1].contiguous()
        prec = _lib.PREC[precision]
        ws, nbytes = _ws(f, B, T, U1, H, Vp, prec)
        stash, stash_bytes = _stash(f, B, T, U1, H, Vp, prec, needs_grad, stash_gib)
        costs = torch.empty(B, dtype=torch.float32, device=f.device)
        sumsq = torch.zeros(B, T, U1, dtype=torch.float32, device=f.device) if want_sumsq else None
        L = _lib.lib()
        with torch.cuda.device(f.device):
            st = L.clasr_joint_rnnt_fwd(
                f.data_ptr(), g.data_ptr(), weight.data_ptr(), bias.data_ptr(), _lib.ptr(labels) if U1 > 1 else 0,
                act_lens.data_ptr(), label_lens.data_ptr(), B, T, U1, H, Vp, int(blank), _lib.ACT[activation], prec,
                float(dropout_p), int(dropout_seed), float(fastemit_lambda), costs.data_ptr(), _lib.ptr(sumsq),
                ws.data_ptr(), nbytes, _lib.ptr(stash), stash_bytes, _lib.stream_ptr(f.device))
        _lib.check(st, "joint_rnnt_fwd")
        ctx.save_for_backward(f, g, weight, bias, labels, act_lens, label_lens, ws)
        ctx.stash = (stash, stash_bytes)
        ctx.args = (int(blank), _lib.ACT[activation], prec, float(fastemit_lambda), float(clamp), nbytes,
                    float(dropout_p), int(dropout_seed))
        if want_sumsq:
            ctx.mark_non_differentiable(sumsq)
            return costs, sumsq
        return costs

    @staticmethod
    def backward(ctx, grad_costs, *unused):
        f, g, weight, bias, labels, act_lens, label_lens, ws = ctx.saved_tensors
        blank, act, prec, fastemit_lambda, clamp, nbytes, dropout_p, dropout_seed = ctx.args
        stash, stash_bytes = ctx.stash
        B, T, H = f.shape
        U1 = g.shape[1]
        Vp = weight.shape[0]
        go = grad_costs.contiguous().float().view(-1)
        d_f = torch.empty_like(f)
        d_g = torch.empty_like(g)
        d_w = torch.empty_like(weight)
        d_b = torch.empty_like(bias)
        L = _lib.lib()
        sbytes = L.clasr_joint_bwd_scratch_bytes(B, T, U1, H, Vp, prec)
        scratch = torch.empty(sbytes, dtype=torch.uint8, device=f.device)
        with torch.cuda.device(f.device):
            st = L.clasr_joint_rnnt_bwd(
                f.data_ptr(), g.data_ptr(), weight.data_ptr(), bias.data_ptr(), _lib.ptr(labels) if U1 > 1 else 0,
                act_lens.data_ptr(), label_lens.data_ptr(), B, T, U1, H, Vp, blank, act, prec, dropout_p, dropout_seed,
                fastemit_lambda, clamp, go.data_ptr(), d_f.data_ptr(), d_g.data_ptr(), d_w.data_ptr(), d_b.data_ptr(),
                ws.data_ptr(), nbytes, scratch.data_ptr(), sbytes, _lib.ptr(stash), stash_bytes,
                _lib.stream_ptr(f.device))
        ctx.stash = (None, 0)   # the logits are dead once dZ exists
        _lib.check(st, "joint_rnnt_bwd")
        return (d_f, d_g, d_w, d_b) + (None,) * 12


def fused_joint_rnnt_loss(f, g, weight, bias, labels, act_lens, label_lens, blank, activation="tanh",
                          precision="fp16x3", fastemit_lambda=0.0, clamp=0.0, dropout_p=0.0, dropout_seed=0,
                          stash_gib=None):
    """Per-sample transducer costs [B].  ``dropout_p`` > 0 applies the joint's Dropout between the activation and the
    output layer inside the kernels (mask keyed by ``dropout_seed``; see ``dropout_mask_reference``).  ``stash_gib``:
    see ``_stash_limit_bytes`` (default: the logits never reach HBM, the backward pass recomputes them)."""
    return _FusedJointRNNT.apply(f, g, weight, bias, labels, act_lens, label_lens, blank, activation, precision,
                                 fastemit_lambda, clamp, False, dropout_p, dropout_seed, stash_gib)


def dropout_mask_reference(act_lens, label_lens, T, U1, H, dropout_p, dropout_seed):
    """Host restatement of the in-kernel dropout mask: bool [B,T,U1,H], True = kept (test / debugging aid).

    A cell's key is its compact tile-row index ``128 * tile_offsets[b] + t * (U_b+1) + u`` with ``tile_offsets`` the
    prefix sums of ``ceil(T_b (U_b+1) / 128)``; one lowbias32 hash per feature pair, 16 bits per feature."""
    import numpy as np

    al = [int(x) for x in act_lens]
    ll = [int(x) for x in label_lens]
    B = len(al)
    thresh = min(65535, int(round(float(np.float32(dropout_p) * np.float32(65536.0)))))
    sa, sb = np.uint32(dropout_seed & 0xFFFFFFFF), np.uint32((dropout_seed >> 32) & 0xFFFFFFFF)
    keep = np.zeros((B, T, U1, H), dtype=bool)
    off = 0
    k = np.arange(H, dtype=np.uint32)
    with np.errstate(over="ignore"):
        for b in range(B):
            Tb, Ub1 = al[b], ll[b] + 1
            t, u = np.meshgrid(np.arange(Tb, dtype=np.uint32), np.arange(Ub1, dtype=np.uint32), indexing="ij")
            row = (np.uint32(off * 128) + t * np.uint32(Ub1) + u)[..., None]
            x = (row * np.uint32(H // 2) + (k >> np.uint32(1)) + sa) ^ sb
            x ^= x >> np.uint32(16); x *= np.uint32(0x7feb352d)
            x ^= x >> np.uint32(15); x *= np.uint32(0x846ca68b)
            x ^= x >> np.uint32(16)
            r16 = np.where((k & np.uint32(1)) == 1, x >> np.uint32(16), x & np.uint32(0xFFFF))
            keep[b, :Tb, :Ub1] = r16 >= thresh
            off += (Tb * Ub1 + 127) // 128
    scale = 65536.0 / (65536.0 - thresh) if thresh else 1.0
    return keep, scale


def fused_joint_forward_stats(f, g, weight, bias, labels, act_lens, label_lens, blank, activation="tanh",
                              precision="fp16x3"):
    """(costs [B], sum_v z^2 [B,T,U+1]) — the second is what MAS needs from the joint logits."""
    return _FusedJointRNNT.apply(f, g, weight, bias, labels, act_lens, label_lens, blank, activation, precision, 0.0,
                                 0.0, True, 0.0, 0)


class _FusedJointSumsq(torch.autograd.Function):
    """sum_v z[b,t,u,v]^2 for every cell selected by (act_lens, label_lens), differentiable w.r.t. f, g, W, b — the
    joint half of the MAS importance objective (reference cl_baseline_mas.py:258-265) without the logits tensor."""

    @staticmethod
    def forward(ctx, f, g, weight, bias, labels, act_lens, label_lens, blank, activation, precision, dropout_p=0.0,
                dropout_seed=0, stash_gib=None):
        _lib.require_cuda(f, "f")
        needs_grad = any(ctx.needs_input_grad[:4])
        f = f.contiguous().float()
        g = g.contiguous().float()
        weight = weight.contiguous().float()
        bias = bias.contiguous().float()
        B, T, H = f.shape
        U1 = g.shape[1]
        Vp = weight.shape[0]
        labels = labels[:, : U1 - 1].contiguous().long() if U1 > 1 else labels.contiguous().long()
        act_lens = act_lens.contiguous().long()
        label_lens = label_lens.contiguous().long()
        prec = _lib.PREC[precision]
        ws, nbytes = _ws(f, B, T, U1, H, Vp, prec)
        stash, stash_bytes = _stash(f, B, T, U1, H, Vp, prec, needs_grad, stash_gib)
        costs = torch.empty(B, dtype=torch.float32, device=f.device)  # by-product of the same pass; not used here
        sumsq = torch.zeros(B, T, U1, dtype=torch.float32, device=f.device)
        L = _lib.lib()
        with torch.cuda.device(f.device):
            st = L.clasr_joint_rnnt_fwd(
                f.data_ptr(), g.data_ptr(), weight.data_ptr(), bias.data_ptr(), _lib.ptr(labels) if U1 > 1 else 0,
                act_lens.data_ptr(), label_lens.data_ptr(), B, T, U1, H, Vp, int(blank), _lib.ACT[activation], prec,
                float(dropout_p), int(dropout_seed), 0.0, costs.data_ptr(), sumsq.data_ptr(), ws.data_ptr(), nbytes,
                _lib.ptr(stash), stash_bytes, _lib.stream_ptr(f.device))
        _lib.check(st, "joint_rnnt_fwd")
        ctx.save_for_backward(f, g, weight, bias, labels, act_lens, label_lens, ws)
        ctx.stash = (stash, stash_bytes)
        ctx.args = (int(blank), _lib.ACT[activation], prec, nbytes, float(dropout_p), int(dropout_seed))
        return sumsq

    @staticmethod
    def backward(ctx, grad_sumsq):
        f, g, weight, bias, labels, act_lens, label_lens, ws = ctx.saved_tensors
        blank, act, prec, nbytes, dropout_p, dropout_seed = ctx.args
        stash, stash_bytes = ctx.stash
        B, T, H = f.shape
        U1 = g.shape[1]
        Vp = weight.shape[0]
        gc = grad_sumsq.contiguous().float()
        d_f, d_g = torch.empty_like(f), torch.empty_like(g)
        d_w, d_b = torch.empty_like(weight), torch.empty_like(bias)
        L = _lib.lib()
        sbytes = L.clasr_joint_bwd_scratch_bytes(B, T, U1, H, Vp, prec)
        scratch = torch.empty(sbytes, dtype=torch.uint8, device=f.device)
        with torch.cuda.device(f.device):
            st = L.clasr_joint_sumsq_bwd(
                f.data_ptr(), g.data_ptr(), weight.data_ptr(), bias.data_ptr(), _lib.ptr(labels) if U1 > 1 else 0,
                act_lens.data_ptr(), label_lens.data_ptr(), B, T, U1, H, Vp, blank, act, prec, dropout_p, dropout_seed,
                gc.data_ptr(), d_f.data_ptr(), d_g.data_ptr(), d_w.data_ptr(), d_b.data_ptr(), ws.data_ptr(), nbytes,
                scratch.data_ptr(), sbytes, _lib.ptr(stash), stash_bytes, _lib.stream_ptr(f.device))
        ctx.stash = (None, 0)
        _lib.check(st, "joint_sumsq_bwd")
        return (d_f, d_g, d_w, d_b) + (None,) * 9


def fused_joint_sumsq(f, g, weight, bias, labels, act_lens, label_lens, blank, activation="tanh", precision="fp16x3",
                      dropout_p=0.0, dropout_seed=0, stash_gib=None):
    """[B,T,U+1] tensor of sum_v z^2 (zero outside the cells selected by the lengths), with autograd."""
    return _FusedJointSumsq.apply(f, g, weight, bias, labels, act_lens, label_lens, blank, activation, precision,
                                  dropout_p, dropout_seed, stash_gib)


class LazySubLogits:
    """Stand-in for one entry of ``RNNTJoint.store_list`` when ``store_sub_logits`` is set on the fused path.

    The only thing the reference's MAS driver does with a stored sub-batch logits tensor is
    ``(x.flatten(end_dim=-2) ** 2).sum(dim=-1).mean()`` (cl_baseline_mas.py:260-262).  This object answers exactly that
    chain from the fused kernel's per-cell sum of squares (so the [b,T',U'+1,V+1] tensor never exists) and falls
    back to materialising the logits for anything else."""

    def __init__(self, sumsq_box: torch.Tensor, vp: int, materialise, squared: bool = False):
        self._sumsq = sumsq_box          # [b, T', U'+1] (or flattened), autograd-connected
        self._vp = vp
        self._materialise = materialise  # () -> logits tensor [b, T', U'+1, V+1]
        self._squared = squared

    @property
    def shape(self):
        return torch.Size(tuple(self._sumsq.shape) + (self._vp,))

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def flatten(self, start_dim: int = 0, end_dim: int = -1):
        if start_dim == 0 and end_dim in (-2, self._sumsq.dim() - 1):
            return LazySubLogits(self._sumsq.reshape(-1), self._vp, lambda: self._materialise().flatten(end_dim=-2),
                                 self._squared)
        return self.materialise().flatten(start_dim, end_dim)

    def __pow__(self, p):
        if p == 2 and not self._squared:
            return LazySubLogits(self._sumsq, self._vp, lambda: self._materialise() ** 2, True)
        return self.materialise() ** p

    def pow(self, p):
        return self.__pow__(p)

    def sum(self, dim=None, **kw):
        if self._squared and dim in (-1, self._sumsq.dim()) and not kw:
            return self._sumsq
        return self.materialise().sum(dim, **kw) if dim is not None else self.materialise().sum(**kw)

    def materialise(self) -> torch.Tensor:
        return self._materialise()

    def __getattr__(self, name):  # any other tensor API: pay for the logits
        return getattr(self.materialise(), name)
