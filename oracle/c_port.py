"""ctypes access to oracle/_build/liboracle.so (lattice.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.check_call(["make", "-C", _HERE])
        _lib = C.CDLL(_SO)
        _lib.oracle_rnnt_cpu.restype = C.c_int
        _lib.oracle_rnnt_cpu.argtypes = [C.c_void_p] * 4 + [C.c_int] * 5 + [C.c_float, C.c_void_p, C.c_void_p]
    return _lib


class _RNNTCpuFn(torch.autograd.Function):
    """costs from log-probs; backward = d/d(log-probs) (autograd then does log_softmax backward, exactly the
    reference's CPU composition, rnnt_pytorch.py:411-437)."""

    @staticmethod
    def forward(ctx, log_probs, labels, act_lens, label_lens, blank, fastemit_lambda):
        lp = log_probs.detach().contiguous()
        B, T, U1, Vp = lp.shape
        costs = torch.empty(B, dtype=torch.float32)
        grads = torch.empty_like(lp)
        rc = lib().oracle_rnnt_cpu(lp.data_ptr(), labels.contiguous().data_ptr(), act_lens.contiguous().data_ptr(),
                                   label_lens.contiguous().data_ptr(), B, T, U1, Vp, int(blank),
                                   float(fastemit_lambda), costs.data_ptr(), grads.data_ptr())
        if rc != 0:
            raise RuntimeError("oracle_rnnt_cpu failed")
        ctx.grads = grads
        return costs

    @staticmethod
    def backward(ctx, go):
        return ctx.grads * go.view(-1, 1, 1, 1), None, None, None, None, None


def rnnt_loss_cpu(logits, labels, act_lens, label_lens, blank, fastemit_lambda=0.0):
    lp = torch.nn.functional.log_softmax(logits.float(), -1)
    return _RNNTCpuFn.apply(lp, labels.long(), act_lens.long(), label_lens.long(), blank, fastemit_lambda)
