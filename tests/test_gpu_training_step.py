"""EncDecHybridRNNTCTCStep.training_step (reference hybrid_rnnt_ctc_models.py:859-930) and the EWC / MAS inner-loop
steps of the drivers (cl_baseline_ewc.py:225-255, cl_baseline_mas.py:212-271) against the fp64 ORACLE composition:
the same encoder / prediction network evaluated in float64 on the CPU, oracle/joint_oracle.py (restated joint + fused
sub-batch loop over oracle/rnnt_oracle.py), oracle/ctc_oracle.py and oracle/cl_oracle.py — not against this package's
own modules.  Tolerances: relative 1e-5 on losses, 1e-4 on gradients (north_star)."""
import copy

import numpy as np
import pytest
import torch

from helpers import rel_err
from indic_cl_asr_b200 import (CTCLoss, ConvASRDecoder, EncDecHybridRNNTCTCStep, RNNTDecoder, RNNTJoint, RNNTLoss, cl)
from indic_cl_asr_b200.hybrid import ewc_backward, mas_importance_backward
from oracle import cl_oracle, ctc_oracle, joint_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
V, H, De, Dp, F = 23, 64, 24, 16, 10
CTC_W = 0.3


class TinyEncoder(torch.nn.Module):
    """Stand-in for preprocessor + Conformer (out of scope): [B,F,Tm] -> [B,De,T'] with 4x striding, `layers` like the
    reference encoder so that utils.freeze_layer applies."""

    def __init__(self):
        super().__init__()
        self.layers = torch.nn.ModuleList([torch.nn.Conv1d(F, De, kernel_size=4, stride=4),
                                           torch.nn.Conv1d(De, De, kernel_size=1)])

    def forward(self, input_signal, input_signal_length):
        x = torch.tanh(self.layers[0](input_signal))
        x = x + torch.tanh(self.layers[1](x))
        return x, torch.div(input_signal_length, 4, rounding_mode="floor")


class _Wer:
    def __init__(self):
        self.calls = []

    def update(self, **kw):
        self.calls.append(sorted(kw))

    def compute(self):
        return torch.tensor(0.25, device=DEV), torch.tensor(1.0, device=DEV), torch.tensor(4.0, device=DEV)

    def reset(self):
        pass


def build(act="tanh", wer=None, ctc_wer=None, monitor_host=True, seed=0):
    torch.manual_seed(seed)
    enc = TinyEncoder()
    dec = RNNTDecoder(prednet=dict(pred_hidden=Dp, pred_rnn_layers=1, dropout=0.0), vocab_size=V)
    joint = RNNTJoint(jointnet=dict(encoder_hidden=De, pred_hidden=Dp, joint_hidden=H, activation=act, dropout=0.0),
                      num_classes=V, fuse_loss_wer=True, fused_batch_size=2)
    head = ConvASRDecoder(feat_in=De, num_classes=V)
    model = EncDecHybridRNNTCTCStep(enc, dec, joint, head, RNNTLoss(num_classes=V), CTCLoss(num_classes=V, zero_infinity=True),
                                    wer=wer, ctc_wer=ctc_wer, ctc_loss_weight=CTC_W, monitor_host=monitor_host)
    return model.to(DEV)


def batch(seed=1, B=5, Tm=64, U=6):
    g = torch.Generator().manual_seed(seed)
    sig = torch.randn(B, F, Tm, generator=g)
    sl = torch.randint(Tm // 2, Tm + 1, (B,), generator=g)
    sl[0] = Tm
    tr = torch.randint(0, V, (B, U), generator=g)
    tl = torch.randint(1, U + 1, (B,), generator=g)
    tl[0] = U
    return sig, sl, tr, tl


def oracle_step(model, b, act, want_sub_logits=False):
    """The same step in float64 on the CPU through the oracle.  Returns (loss, rnnt, ctc, log_probs, {name: grad})
    (+ the sub-batch logits and CTC logits with their graphs when asked)."""
    sig, sl, tr, tl = b
    enc64 = copy.deepcopy(model.encoder).cpu().double()
    dec64 = copy.deepcopy(model.decoder).cpu().double()
    p = {"enc.weight": model.joint.enc.weight, "enc.bias": model.joint.enc.bias, "pred.weight": model.joint.pred.weight,
         "pred.bias": model.joint.pred.bias, "out.weight": model.joint.joint_net[-1].weight,
         "out.bias": model.joint.joint_net[-1].bias}
    p = {k: v.detach().cpu().double().requires_grad_(True) for k, v in p.items()}
    cw = model.ctc_decoder.decoder_layers[0].weight.detach().cpu().double().requires_grad_(True)
    cb = model.ctc_decoder.decoder_layers[0].bias.detach().cpu().double().requires_grad_(True)
    encoded, el = enc64(sig.double(), sl)
    g, _, _ = dec64(targets=tr, target_length=tl)
    out = joint_oracle.fused_joint_loss(encoded, g, el, tr, tl, p, act, V, 2, "mean_batch", return_sub_logits=want_sub_logits)
    o_rnnt, subs = out if want_sub_logits else (out, None)
    lp, z_ctc = joint_oracle.ctc_head(encoded, cw, cb)

    class _Ctc(torch.autograd.Function):
        @staticmethod
        def forward(ctx, lp_):
            nll, gr = ctc_oracle.ctc_loss_and_grad(lp_.detach().numpy(), tr.numpy(), el.numpy(), tl.numpy(), V, True)
            ctx.g = torch.from_numpy(gr)
            return torch.from_numpy(nll)

        @staticmethod
        def backward(ctx, go):
            return ctx.g * go.view(-1, 1, 1)

    o_ctc = _Ctc.apply(lp).mean()
    o_loss = (1 - CTC_W) * o_rnnt + CTC_W * o_ctc
    names = {}
    for k, t in enc64.named_parameters():
        names["encoder." + k] = t
    for k, t in dec64.named_parameters():
        names["decoder." + k] = t
    last = len(model.joint.joint_net) - 1
    names.update({"joint.enc.weight": p["enc.weight"], "joint.enc.bias": p["enc.bias"], "joint.pred.weight": p["pred.weight"],
                  "joint.pred.bias": p["pred.bias"], f"joint.joint_net.{last}.weight": p["out.weight"],
                  f"joint.joint_net.{last}.bias": p["out.bias"], "ctc_decoder.decoder_layers.0.weight": cw,
                  "ctc_decoder.decoder_layers.0.bias": cb})
    return o_loss, o_rnnt, o_ctc, lp, names, subs, z_ctc


def trainable(model):
    return {n: p for n, p in model.named_parameters() if p.requires_grad}


@pytest.mark.parametrize("act", ["tanh", "relu"])
def test_training_step_vs_oracle(act):
    wer, ctc_wer = _Wer(), _Wer()
    model = build(act, wer=wer, ctc_wer=ctc_wer)
    b = batch()
    dev_b = [x.to(DEV) for x in b]
    loss, monitor, log_probs = model.training_step(dev_b, None, return_probs=True)
    loss.backward()
    torch.cuda.synchronize()
    o_loss, o_rnnt, o_ctc, o_lp, names, _, _ = oracle_step(model, b, act)
    o_loss.backward()
    assert sorted(monitor) == ["train_ctc_loss", "train_loss", "train_rnnt_loss", "training_batch_wer",
                               "training_batch_wer_ctc"]                      # hybrid_rnnt_ctc_models.py:889-920
    assert isinstance(monitor["train_loss"], float) and isinstance(monitor["training_batch_wer_ctc"], float)
    assert abs(monitor["train_rnnt_loss"] - o_rnnt.item()) <= 1e-5 * abs(o_rnnt.item())
    assert abs(monitor["train_ctc_loss"] - o_ctc.item()) <= 1e-5 * abs(o_ctc.item())
    assert abs(loss.item() - o_loss.item()) <= 1e-5 * abs(o_loss.item())
    assert monitor["training_batch_wer_ctc"] == 0.25 and float(monitor["training_batch_wer"]) == 0.25
    assert len(wer.calls) == 3 and ctc_wer.calls == [["predictions", "predictions_lengths", "targets", "targets_lengths"]]
    el = torch.div(b[1], 4, rounding_mode="floor")
    for i in range(len(el)):   # rows past an utterance's length are padding (log-softmax of whatever the encoder emitted)
        assert np.allclose(log_probs[i, : int(el[i])].detach().cpu().numpy(), o_lp[i, : int(el[i])].detach().numpy(), atol=2e-5)
    got = trainable(model)
    assert sorted(got) == sorted(names)
    for n, p in got.items():
        assert rel_err(p.grad.cpu().numpy(), names[n].grad.numpy()) <= 1e-4, n


def test_training_step_device_monitor_never_syncs_and_lang_ids():
    model = build("relu", monitor_host=False)
    out = model.training_step([x.to(DEV) for x in batch(seed=4)], None)
    assert len(out) == 2
    loss, monitor = out
    assert monitor["training_batch_wer"] is None and monitor["training_batch_wer_ctc"] is None
    assert all(monitor[k].is_cuda for k in ("train_rnnt_loss", "train_ctc_loss", "train_loss"))
    assert torch.allclose(monitor["train_loss"], (1 - CTC_W) * monitor["train_rnnt_loss"] + CTC_W * monitor["train_ctc_loss"])


def test_ewc_step_and_fisher_accumulation_vs_oracle():
    """cl_baseline_ewc.py:228-255: penalty gradient pre-loaded, backward on top, then (importance epoch)
    F += mean(loss) * grad^2 — against cl_oracle on the oracle's fp64 gradients."""
    model = build("tanh")
    cl.freeze_layer(model, 0)          # utils.py:246-263: encoder.layers[0] frozen, everything else trainable
    b = batch(seed=7)
    star = cl.get_params_clone(model)
    star.flat.add_(0.01 * torch.randn_like(star.flat))
    main_fish = cl.get_zero_params(model, DEV)
    main_fish.flat.uniform_(0.0, 1.0)
    cfg = {"cl_config": {"e_lambda": 10.0}}
    loss, _ = model.training_step([x.to(DEV) for x in b], None)
    avg = ewc_backward(model, loss, cfg, main_fish, star)
    torch.cuda.synchronize()
    o_loss, _, _, _, names, _, _ = oracle_step(model, b, "tanh")
    o_loss.backward()
    got = trainable(model)
    assert "encoder.layers.0.weight" not in got and "encoder.layers.1.weight" in got
    theta = {n: p.detach().cpu().double() for n, p in got.items()}
    o_pen, o_avg = cl_oracle.get_penalty_grads(10.0, {n: main_fish[n].cpu().double() for n in got}, theta,
                                               {n: star[n].cpu().double() for n in got})
    assert abs(avg.item() - o_avg) <= 1e-5 * abs(o_avg)
    for n, p in got.items():
        ref = names[n].grad.numpy() + o_pen[n].numpy()
        assert rel_err(p.grad.cpu().numpy(), ref) <= 1e-4, n
    # importance epoch (no penalty, no optimiser step): Fisher accumulation weighted by the batch loss
    fish = cl.get_zero_params(model, DEV)
    fp = cl.flat_params(model)
    fp.bind_grads(zero=True)
    loss2, _ = model.training_step([x.to(DEV) for x in b], None)
    loss2.backward()
    cl.fisher_accumulate(fish, fp.grad_dict(), loss2)
    torch.cuda.synchronize()
    o_f = {n: torch.zeros_like(t) for n, t in theta.items()}
    cl_oracle.fisher_accumulate(o_f, {n: names[n].grad for n in got}, o_loss)
    for n in got:
        assert rel_err(fish[n].cpu().numpy(), o_f[n].numpy()) <= 2e-4, n     # squares of 1e-4-accurate gradients


@pytest.mark.parametrize("act", ["tanh", "relu"])
def test_mas_importance_step_vs_oracle(act):
    """cl_baseline_mas.py:212-221, 257-271: hook flags on, training_step, objective over the stored sub-batch logits and
    the CTC logits, backward, Omega += |grad| — against joint_oracle.mas_objective on the oracle's fp64 logits."""
    model = build(act)
    b = batch(seed=9)
    model.joint.store_sub_enc, model.joint.store_sub_logits, model.joint.detach_sub_enc = False, True, False
    model.ctc_decoder.return_logits_ = True
    imp = cl.get_zero_params(model, DEV)
    cl.flat_params(model).bind_grads(zero=True)
    model.training_step([x.to(DEV) for x in b], None)
    obj = mas_importance_backward(model, model.joint, model.ctc_decoder, imp, mas_ctx=0.3)
    torch.cuda.synchronize()
    _, _, _, _, names, subs, z_ctc = oracle_step(model, b, act, want_sub_logits=True)
    o_obj = joint_oracle.mas_objective(subs, z_ctc, 0.3)
    o_obj.backward()
    assert abs(obj.item() - o_obj.item()) <= 1e-5 * abs(o_obj.item())
    for n in trainable(model):
        assert rel_err(imp[n].cpu().numpy(), names[n].grad.abs().numpy()) <= 1e-4, n
