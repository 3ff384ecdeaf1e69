"""RNNTDecoder (prediction network) with the reference's interface — the producer of the joint's ``g`` operand.

Mirrors NeMo/nemo/collections/asr/modules/rnnt.py:524-1173 (``RNNTDecoder``): ``Embedding(vocab+1, H, padding_idx=blank)``
-> prepend an all-zero start-of-sequence step (:769-775) -> ``LSTMDropout`` (common/parts/rnn.py:151-236: torch.nn.LSTM
with forget-gate-bias / chrono initialisation and a weight scale).  Parameter names (``prediction.embed.weight``,
``prediction.dec_rnn.lstm.weight_ih_l0`` ...) are the reference's, so state_dicts interchange.

Scope: the recurrence itself stays on cuDNN through ``torch.nn.LSTM``.  At the benchmark
shape it is U+1 = 101 dependent steps of a [B,H] x [H,4H] GEMM (105 MFLOP each at B=32, H=640) — latency-bound
library work that is not one of the four kernels this library replaces.  What this module adds on top of the library
call is the layout contract with the fused joint: the LSTM's time-major output [U+1,B,H] is handed over as the
[B,D,U+1] view the reference returns (:667-681), which ``RNNTJoint.forward`` turns back into [B,U+1,H] for the
``pred`` projection without an intermediate copy of its own.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import torch

from .._typecheck import kwargs_only

__all__ = ["RNNTDecoder", "LSTMDropout", "label_collate"]


def label_collate(labels, device=None) -> torch.Tensor:
    """common/parts/rnn.py:536-560: tensors pass through as int64, lists are zero-padded to [B, max_len]."""
    if isinstance(labels, torch.Tensor):
        return labels.type(torch.int64)
    if not isinstance(labels, (list, tuple)):
        raise ValueError(f"`labels` should be a list or tensor not {type(labels)}")
    max_len = max(len(label) for label in labels)
    out = torch.zeros((len(labels), max_len), dtype=torch.int64)
    for e, l in enumerate(labels):
        out[e, : len(l)] = torch.as_tensor(l, dtype=torch.int64)
    return out.to(device) if device is not None else out


class LSTMDropout(torch.nn.Module):
    """common/parts/rnn.py:151-236."""

    def __init__(self, input_size: int, hidden_size: int, num_layers: int, dropout: Optional[float],
                 forget_gate_bias: Optional[float], t_max: Optional[int] = None, weights_init_scale: float = 1.0,
                 hidden_hidden_bias_scale: float = 0.0, proj_size: int = 0):
        super().__init__()
        self.lstm = torch.nn.LSTM(input_size=input_size, hidden_size=hidden_size, num_layers=num_layers,
                                  dropout=dropout or 0.0, proj_size=proj_size)
        with torch.no_grad():
            if t_max is not None:  # chrono initialisation (:195-207)
                for name, p in self.lstm.named_parameters():
                    if "bias" in name:
                        h = p.nelement() // 4
                        p.fill_(0)
                        p[h:2 * h] = torch.log(torch.nn.init.uniform_(p[0:h].clone(), 1, t_max - 1))
                        p[0:h] = -p[h:2 * h]
            elif forget_gate_bias is not None:  # (:209-216)
                for name, p in self.lstm.named_parameters():
                    if "bias_ih" in name:
                        p[hidden_size:2 * hidden_size].fill_(forget_gate_bias)
                    if "bias_hh" in name:
                        p[hidden_size:2 * hidden_size] *= float(hidden_hidden_bias_scale)
            self.dropout = torch.nn.Dropout(dropout) if dropout else None
            for name, p in self.named_parameters():
                if "weight" in name or "bias" in name:
                    p *= float(weights_init_scale)

    def forward(self, x: torch.Tensor, h=None):
        x, h = self.lstm(x, h)
        if self.dropout:
            x = self.dropout(x)
        return x, h


class RNNTDecoder(torch.nn.Module):
    def __init__(self, prednet: Dict[str, Any], vocab_size: int, normalization_mode: Optional[str] = None,
                 random_state_sampling: bool = False, blank_as_pad: bool = True, multisoftmax=False,
                 language_masks=None):
        super().__init__()
        if normalization_mode is not None:
            raise NotImplementedError("RNNTDecoder: batch / layer normalised RNN stacks (common/parts/rnn.py:88-130) are "
                                      "not part of the shipped checkpoint's prediction network (norm = None)")
        self.pred_hidden = prednet["pred_hidden"]
        self.pred_rnn_layers = prednet["pred_rnn_layers"]
        self.blank_idx = vocab_size
        self.vocab_size = vocab_size
        self.blank_as_pad = blank_as_pad
        self.random_state_sampling = random_state_sampling
        self.multisoftmax = multisoftmax
        self.language_masks = language_masks
        rnn_hidden = prednet.get("rnn_hidden_size", -1)
        if blank_as_pad:
            embed = torch.nn.Embedding(vocab_size + 1, self.pred_hidden, padding_idx=self.blank_idx)
        else:
            embed = torch.nn.Embedding(vocab_size, self.pred_hidden)
        self.prediction = torch.nn.ModuleDict({
            "embed": embed,
            "dec_rnn": LSTMDropout(
                input_size=self.pred_hidden, hidden_size=rnn_hidden if rnn_hidden > 0 else self.pred_hidden,
                num_layers=self.pred_rnn_layers, dropout=prednet.get("dropout", 0.0),
                forget_gate_bias=prednet.get("forget_gate_bias", 1.0), t_max=prednet.get("t_max", None),
                weights_init_scale=prednet.get("weights_init_scale", 1.0),
                hidden_hidden_bias_scale=prednet.get("hidden_hidden_bias_scale", 0.0),
                proj_size=self.pred_hidden if self.pred_hidden < rnn_hidden else 0),
        })
        self._rnnt_export = False

    def is_adapter_available(self) -> bool:
        return False

    @kwargs_only
    def forward(self, targets, target_length, states=None):
        """-> (g [B, D, U+1], target_length, states)   (reference :667-681)."""
        y = label_collate(targets)
        g, states = self.predict(y, state=states, add_sos=not self._rnnt_export)   # [B, U+1, D]
        return g.transpose(1, 2), target_length, states

    def predict(self, y: Optional[torch.Tensor] = None, state: Optional[List[torch.Tensor]] = None,
                add_sos: bool = True, batch_size: Optional[int] = None) -> Tuple[torch.Tensor, List[torch.Tensor]]:
        """Reference :683-784: embed (or a zero step when ``y`` is None), optional zero SOS step, LSTM."""
        _p = next(self.parameters())
        device, dtype = _p.device, _p.dtype
        if y is not None:
            if y.device != device:
                y = y.to(device)
            y = self.prediction["embed"](y)
        else:
            B = batch_size if batch_size is not None else (1 if state is None else state[0].size(1))
            y = torch.zeros((B, 1, self.pred_hidden), device=device, dtype=dtype)
        if add_sos:
            B, U, H = y.shape
            y = torch.cat([torch.zeros((B, 1, H), device=y.device, dtype=y.dtype), y], dim=1).contiguous()
        if state is None and self.random_state_sampling and self.training:
            state = self.initialize_state(y)
        g, hid = self.prediction["dec_rnn"](y.transpose(0, 1), state)   # time-major in, [U+1, B, H] out
        return g.transpose(0, 1), hid

    def initialize_state(self, y: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Reference :845-872: zeros, or N(0,1) samples under random_state_sampling in training mode."""
        batch = y.size(0)
        shape = (self.pred_rnn_layers, batch, self.pred_hidden)
        if self.random_state_sampling and self.training:
            return (torch.randn(*shape, dtype=y.dtype, device=y.device), torch.randn(*shape, dtype=y.dtype, device=y.device))
        return (torch.zeros(*shape, dtype=y.dtype, device=y.device), torch.zeros(*shape, dtype=y.dtype, device=y.device))
