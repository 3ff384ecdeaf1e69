"""The loss half of ``EncDecHybridRNNTCTCModel.training_step`` and the EWC / MAS inner-loop steps of the drivers,
composed from the B200 modules without the reference's host synchronisations.

Reference: NeMo/nemo/collections/asr/models/hybrid_rnnt_ctc_models.py:868-902 (joint -> RNNT loss, CTC head -> CTC
loss, ``(1-w) * rnnt + w * ctc``) — there every step also runs six ``gc.collect(); torch.cuda.empty_cache()`` pairs
(:862-924), four ``.item()`` host syncs for the monitor dict (:899,900,912,920) and a greedy decode for WER
(``compute_wer=True``, :875); cl_baseline_ewc.py:228-255 and cl_baseline_mas.py:231-240,257-271 for the regulariser
steps.  Here the monitor values stay device tensors, the WER hook is only called when asked for, and the CTC branch
(head GEMM, log-softmax, CTC lattice) is enqueued on a side stream so that its small latency-bound kernels overlap
the transducer lattice instead of queueing behind it.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import cl

__all__ = ["HybridRNNTCTCLoss", "EncDecHybridRNNTCTCStep", "ewc_backward", "mas_importance_backward", "silence_side_stream_grad_warning"]


def silence_side_stream_grad_warning() -> None:
    """With ``overlap_ctc`` the CTC head's gradients are accumulated on the side stream; torch (>= 2.9) warns about
    that once per parameter although the autograd engine synchronises the streams correctly.  The switch is
    PROCESS-WIDE, so it is the application's call (bench.py makes it), not a side effect of building a module."""
    warn_off = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
    if warn_off is not None:
        warn_off(False)


class HybridRNNTCTCLoss(torch.nn.Module):
    def __init__(self, joint, ctc_decoder, ctc_loss, ctc_loss_weight: float = 0.3, overlap_ctc: bool = True):
        super().__init__()
        self.joint = joint
        self.ctc_decoder = ctc_decoder
        self.ctc_loss = ctc_loss
        self.ctc_loss_weight = float(ctc_loss_weight)   # cfg.aux_ctc.ctc_loss_weight (:233)
        self.overlap_ctc = overlap_ctc
        self.return_log_probs = False   # also hand the CTC log-probs back in the monitor (training_step's return_probs)
        self._side: Dict[torch.device, torch.cuda.Stream] = {}

    def _side_stream(self, device) -> torch.cuda.Stream:
        if device not in self._side:
            self._side[device] = torch.cuda.Stream(device=device)
        return self._side[device]

    def forward(self, encoded: torch.Tensor, encoded_len: torch.Tensor, decoder: torch.Tensor,
                transcript: torch.Tensor, transcript_len: torch.Tensor, language_ids=None,
                compute_wer: bool = False) -> Tuple[torch.Tensor, Dict[str, Optional[torch.Tensor]]]:
        """``encoded`` [B,D,T'] / ``decoder`` [B,D,U+1] in NeMo layout.  Returns ``(loss, monitor)`` where the monitor
        values are 0-d DEVICE tensors (call ``.item()`` only when something is actually logged)."""
        if not encoded.is_cuda:
            raise RuntimeError("HybridRNNTCTCLoss: CUDA tensors only (no CPU path in indic_cl_asr_b200)")
        cur = torch.cuda.current_stream(encoded.device)
        kw = {} if language_ids is None else {"language_ids": language_ids}

        kept = {}

        def ctc_branch():
            log_probs = self.ctc_decoder(encoder_output=encoded, **kw)                                  # :894
            if self.return_log_probs:
                kept["log_probs"] = log_probs
            return self.ctc_loss(log_probs=log_probs, targets=transcript, input_lengths=encoded_len,
                                 target_lengths=transcript_len)                                         # :896-898

        if self.overlap_ctc:
            side = self._side_stream(encoded.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                loss_ctc = ctc_branch()
            # no record_stream(): the inputs stay referenced by the caller / the autograd graph until backward has run,
            # the main stream joins the side stream below, and autograd joins all streams at the end of backward;
            # record_stream() made the caching allocator defer block reuse and fall into cudaMalloc/cudaFree churn
            # (end-to-end step 13.5 -> 17-20 ms, erratic)
        loss_rnnt, wer, _, _ = self.joint(encoder_outputs=encoded, decoder_outputs=decoder,
                                          encoder_lengths=encoded_len, transcripts=transcript,
                                          transcript_lengths=transcript_len, compute_wer=compute_wer, **kw)  # :880-888
        if self.overlap_ctc:
            cur.wait_stream(side)
        else:
            loss_ctc = ctc_branch()
        w = self.ctc_loss_weight
        loss = (1 - w) * loss_rnnt + w * loss_ctc                                                       # :902
        monitor = {"train_rnnt_loss": loss_rnnt.detach(), "train_ctc_loss": loss_ctc.detach(),
                   "train_loss": loss.detach(), "training_batch_wer": wer}
        monitor.update(kept)
        return loss, monitor


class EncDecHybridRNNTCTCStep(torch.nn.Module):
    """``EncDecHybridRNNTCTCModel.training_step(batch, lang_ids, return_probs=False)`` work-alike
    (reference hybrid_rnnt_ctc_models.py:859-930) over the B200 modules.

    Same attribute contract as the reference model — ``encoder``, ``decoder``, ``joint``, ``ctc_decoder``, ``loss``,
    ``ctc_loss``, ``wer``, ``ctc_wer``, ``ctc_loss_weight`` (so ``utils.freeze_layer`` and the drivers' hook flags
    ``model.joint.store_sub_logits`` / ``model.ctc_decoder.return_logits_`` work unchanged) — same call signature, same
    return value ``(loss, monitor[, log_probs])`` and the same five monitor keys.  What is NOT reproduced is the
    reference's per-step overhead: six ``gc.collect(); torch.cuda.empty_cache()`` pairs (:862-924) and four separate
    ``.item()`` host syncs (:899,900,912,920).  With ``monitor_host=True`` (default, the reference's python floats) all
    monitor scalars cross to the host in ONE copy at the end of the step; with ``monitor_host=False`` they stay 0-d
    device tensors and the step never synchronises.

    ``encoder`` is any module mapping ``(input_signal=, input_signal_length=) -> (encoded [B,D,T'], encoded_len)`` — the
    preprocessor + SpecAugment + Conformer stack is upstream of the path this library accelerates.
    ``wer`` / ``ctc_wer`` (torchmetrics-style ``update / compute / reset``) are called like the reference does
    (``compute_wer = True``, :875, :903-911) when given; pass ``None`` to keep greedy decoding off the training step.
    """

    def __init__(self, encoder, decoder, joint, ctc_decoder, loss, ctc_loss, wer=None, ctc_wer=None,
                 ctc_loss_weight: float = 0.3, overlap_ctc: bool = True, monitor_host: bool = True):
        super().__init__()
        self.encoder = encoder
        self.decoder = decoder
        self.joint = joint
        self.ctc_decoder = ctc_decoder
        self.loss = loss
        self.ctc_loss = ctc_loss
        self.wer = wer
        self.ctc_wer = ctc_wer
        self.ctc_loss_weight = float(ctc_loss_weight)
        self.monitor_host = monitor_host
        # EncDecRNNTModel.__init__ (rnnt_models.py:120-124): the fused joint owns the loss and the WER metric
        if self.joint.fuse_loss_wer:
            self.joint.set_loss(self.loss)
            self.joint.set_wer(self.wer if self.wer is not None else _NoWer())
        self._step = HybridRNNTCTCLoss(joint, ctc_decoder, ctc_loss, ctc_loss_weight=ctc_loss_weight,
                                       overlap_ctc=overlap_ctc)

    def forward(self, input_signal=None, input_signal_length=None):
        """Encoder-only forward, as EncDecRNNTModel.forward (rnnt_models.py:606-655)."""
        return self.encoder(input_signal=input_signal, input_signal_length=input_signal_length)

    def training_step(self, batch, lang_ids, return_probs: bool = False):
        signal, signal_len, transcript, transcript_len = batch
        language_ids = lang_ids
        encoded, encoded_len = self.forward(input_signal=signal, input_signal_length=signal_len)          # :866
        decoder, _, _ = self.decoder(targets=transcript, target_length=transcript_len)                    # :871
        compute_wer = self.wer is not None                                                                # :875
        self._step.return_log_probs = return_probs or self.ctc_wer is not None
        loss_value, mon = self._step(encoded, encoded_len, decoder, transcript, transcript_len,
                                     language_ids=language_ids, compute_wer=compute_wer)                  # :880-902
        log_probs = mon.pop("log_probs", None)
        ctc_wer = None
        if self.ctc_wer is not None:                                                                      # :903-911
            kw = {} if language_ids is None else {"lang_ids": language_ids}
            self.ctc_wer.update(predictions=log_probs.detach(), targets=transcript, targets_lengths=transcript_len,
                                predictions_lengths=encoded_len, **kw)
            ctc_wer, _, _ = self.ctc_wer.compute()
            self.ctc_wer.reset()
        monitor = {"training_batch_wer": mon["training_batch_wer"], "train_rnnt_loss": mon["train_rnnt_loss"],
                   "train_ctc_loss": mon["train_ctc_loss"], "training_batch_wer_ctc": ctc_wer,
                   "train_loss": mon["train_loss"]}
        if self.monitor_host:
            monitor = _monitor_to_host(monitor)
        if return_probs:
            return loss_value, monitor, log_probs
        return loss_value, monitor


class _NoWer:
    """Stands in for the WER metric when none is attached (the fused joint insists on one, modules/rnnt.py:1407-1410);
    only ever reached if a caller passes ``compute_wer=True`` by hand."""

    def update(self, **kw):
        pass

    def compute(self):
        return None, None, None

    def reset(self):
        pass


def _monitor_to_host(monitor):
    """The reference's monitor holds python floats (``.item()`` per key); here: ONE device->host copy for all keys."""
    keys = [k for k, v in monitor.items() if isinstance(v, torch.Tensor) and v.is_cuda and v.numel() == 1]
    if keys:
        vals = torch.stack([monitor[k].detach().reshape(()).to(torch.float32) for k in keys]).tolist()
        monitor = dict(monitor)
        for k, v in zip(keys, vals):
            if k in ("train_rnnt_loss", "train_ctc_loss", "train_loss", "training_batch_wer_ctc"):
                monitor[k] = v               # floats in the reference (:899,900,912,920)
    return monitor


def ewc_backward(model, loss: torch.Tensor, config, main_fish, checkpoint) -> Optional[torch.Tensor]:
    """cl_baseline_ewc.py:228-240: pre-load ``2 * e_lambda * F * (theta - theta*)`` into the gradients, then
    back-propagate the loss on top.  One fused sweep writes the penalty straight into the model's flat gradient buffer;
    returns ``penalty_avg`` as a device tensor (the reference's ``.item()`` at :81 is left to the caller)."""
    avg = None
    if checkpoint is not None:
        fp = cl.flat_params(model)
        fp.bind_grads(zero=False)
        _, avg = cl.get_penalty_grads_async(config, main_fish, cl.get_params(model), checkpoint, out=fp.grad)
    loss.backward()
    return avg


def mas_importance_backward(model, joint, ctc_decoder, importance, mas_ctx: float) -> torch.Tensor:
    """cl_baseline_mas.py:257-270 after a ``training_step`` run with ``joint.store_sub_logits = True`` and
    ``ctc_decoder.return_logits_ = True``: objective = (1-ctx) * mean_s mean_cells sum_v z_s^2 + ctx * mean sum_v z_ctc^2,
    backward, ``importance += |grad|`` (one sweep over the flat buffers).  Returns the objective (device tensor)."""
    decoder_logits = (ctc_decoder.decoder_logits.flatten(end_dim=-2) ** 2).sum(dim=-1).mean()
    rnn_logits = 0
    for s in joint.store_list:
        rnn_logits = rnn_logits + (s.flatten(end_dim=-2) ** 2).sum(dim=-1).mean()
    rnn_logits = rnn_logits / len(joint.store_list)
    objective = rnn_logits * (1 - mas_ctx) + decoder_logits * mas_ctx
    objective.backward()
    cl.mas_accumulate(importance, model)
    return objective.detach()
