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

__all__ = ["HybridRNNTCTCLoss", "ewc_backward", "mas_importance_backward"]


class HybridRNNTCTCLoss(torch.nn.Module):
    def __init__(self, joint, ctc_decoder, ctc_loss, ctc_loss_weight: float = 0.3, overlap_ctc: bool = True):
        super().__init__()
        self.joint = joint
        self.ctc_decoder = ctc_decoder
        self.ctc_loss = ctc_loss
        self.ctc_loss_weight = float(ctc_loss_weight)   # cfg.aux_ctc.ctc_loss_weight (:233)
        self.overlap_ctc = overlap_ctc
        self._side: Dict[torch.device, torch.cuda.Stream] = {}
        if overlap_ctc:
            # parameters shared by both branches (none here) or created earlier accumulate on their own stream;
            # the engine synchronises correctly, the warning is only about graph capture
            warn_off = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
            if warn_off is not None:
                warn_off(False)

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

        def ctc_branch():
            log_probs = self.ctc_decoder(encoder_output=encoded, **kw)                                  # :894
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
        return loss, monitor


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
