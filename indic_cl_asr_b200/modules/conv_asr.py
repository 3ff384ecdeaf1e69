"""ConvASRDecoder (the CTC head) with the reference's interface
(NeMo/nemo/collections/asr/modules/conv_asr.py:402-528).

Conv1d(k=1) -> optional per-language column select -> (``decoder_logits`` hook for MAS) -> log_softmax.
The reference rebuilds a [B,T,C] boolean mask on the host for every call and runs ``masked_select``
(:471-484); the mask is constant along B (one language per batch in the drivers) and T, so the same result
is an ``index_select`` with a device-resident column index that is built once per language.
log_softmax (+ its backward) runs in this library's row kernels.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .. import _lib
from .._typecheck import kwargs_only

__all__ = ["ConvASRDecoder", "log_softmax_rows"]


class _LogSoftmaxRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _lib.require_cuda(x, "logits")
        x = x.contiguous()
        y = torch.empty_like(x)
        cols = x.shape[-1]
        rows = x.numel() // cols
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().clasr_log_softmax_fwd(x.data_ptr(), y.data_ptr(), rows, cols,
                                                        _lib.stream_ptr(x.device)), "log_softmax_fwd")
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(y)
        cols = y.shape[-1]
        rows = y.numel() // cols
        with torch.cuda.device(y.device):
            _lib.check(_lib.lib().clasr_log_softmax_bwd(y.data_ptr(), dy.data_ptr(), dx.data_ptr(), rows, cols,
                                                        _lib.stream_ptr(y.device)), "log_softmax_bwd")
        return dx


def log_softmax_rows(x: torch.Tensor) -> torch.Tensor:
    if x.dtype != torch.float32:
        x = x.float()
    return _LogSoftmaxRows.apply(x)


class ConvASRDecoder(torch.nn.Module):
    def __init__(self, feat_in, num_classes, init_mode="xavier_uniform", vocabulary=None, multisoftmax=False,
                 language_masks=None):
        super().__init__()
        if vocabulary is None and num_classes < 0:
            raise ValueError("Neither of the vocabulary and num_classes are set! At least one of them need to be set.")
        if num_classes <= 0:
            num_classes = len(vocabulary)
        if vocabulary is not None:
            if num_classes != len(vocabulary):
                raise ValueError(
                    f"If vocabulary is specified, it's length should be equal to the num_classes. "
                    f"Instead got: num_classes={num_classes} and len(vocabulary)={len(vocabulary)}")
            self._vocabulary = vocabulary
        self._feat_in = feat_in
        self._num_classes = num_classes + 1  # +1 blank
        self.decoder_layers = torch.nn.Sequential(
            torch.nn.Conv1d(self._feat_in, self._num_classes, kernel_size=1, bias=True))
        if init_mode == "xavier_uniform":  # parts/submodules/jasper.py init_weights
            torch.nn.init.xavier_uniform_(self.decoder_layers[0].weight, gain=1.0)
        self.temperature = 1.0
        self.multisoftmax = multisoftmax
        self.language_masks = language_masks
        self.return_logits_ = False
        self._column_index: Dict[str, torch.Tensor] = {}

    def is_adapter_available(self) -> bool:
        return False

    @property
    def vocabulary(self):
        return self._vocabulary

    @property
    def num_classes_with_blank(self):
        return self._num_classes

    def _columns(self, lang, device) -> torch.Tensor:
        key = f"{lang}@{device}"
        if key not in self._column_index:
            mask = torch.as_tensor(self.language_masks[lang], dtype=torch.bool)
            self._column_index[key] = torch.nonzero(mask).flatten().to(device)
        return self._column_index[key]

    @kwargs_only
    def forward(self, encoder_output, language_ids=None):
        """Conv1d(k=1) -> per-language column select -> log_softmax (reference modules/conv_asr.py:458-490).

        A kernel-size-1 convolution over [B,D,T] is the GEMM [B*T,D] x [C,D]^T: it runs on the tcgen05 GEMM
        (``linear_x3``) and lands directly in the [B,T,C] layout the reference reaches by transposing.  With one
        language per batch (what the drivers pass, cl_baseline_ewc.py:225) only that language's rows of the weight
        are multiplied (device-resident column index instead of the reference's per-call host-built bool mask and
        the [B,T,n_lang*256+1] intermediate, conv_asr.py:471-484); gradients scatter back into the full weight."""
        from ..linear import linear_x3
        if not encoder_output.is_cuda:
            raise RuntimeError("ConvASRDecoder: CUDA tensors only (no CPU path in indic_cl_asr_b200)")
        conv = self.decoder_layers[0]
        weight, bias = conv.weight.squeeze(-1), conv.bias  # [C,D], [C]
        x = encoder_output.transpose(1, 2)                   # [B,T,D]
        if language_ids is not None and len(set(language_ids)) == 1:
            cols = self._columns(language_ids[0], x.device)
            out = linear_x3(x, weight.index_select(0, cols), bias.index_select(0, cols))
        else:
            out = linear_x3(x, weight, bias)
            if language_ids is not None:  # mixed-language batch: per-sample select (all languages have equal width)
                out = torch.stack([o.index_select(-1, self._columns(l, out.device))
                                   for o, l in zip(out, language_ids)])
        if self.temperature != 1.0:
            out = out / self.temperature
        if self.return_logits_:
            self.decoder_logits = out.clone()
        return log_softmax_rows(out)
