"""Shared builders for the parity tests and __graft_entry__.smoke()."""
from __future__ import annotations

import numpy as np
import torch


def rel_err(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max() / den)


def synth_batch(B, T, U, V, De, Dp, seed=0, ragged=True, device="cpu"):
    """Synthetic inputs in NeMo layouts: enc [B,De,T], dec [B,Dp,U+1], transcripts [B,U], lengths."""
    g = torch.Generator().manual_seed(seed)
    enc = torch.randn(B, De, T, generator=g)
    dec = torch.randn(B, Dp, U + 1, generator=g)
    tr = torch.randint(0, V, (B, U), generator=g)
    if ragged:
        el = torch.randint(max(1, T // 2), T + 1, (B,), generator=g)
        tl = torch.randint(max(0, U // 2), U + 1, (B,), generator=g)
        el[0], tl[0] = T, U
    else:
        el, tl = torch.full((B,), T), torch.full((B,), U)
    return tuple(x.to(device) for x in (enc, dec, tr, el, tl))


def joint_params(joint):
    """Oracle-side parameter dict of an RNNTJoint (single-language head)."""
    lin = joint.joint_net[-1]
    return {"enc.weight": joint.enc.weight, "enc.bias": joint.enc.bias, "pred.weight": joint.pred.weight,
            "pred.bias": joint.pred.bias, "out.weight": lin.weight, "out.bias": lin.bias}


def run_step_and_oracle(device="cuda:0", B=3, T=20, U=7, V=40, H=64, De=32, Dp=32, activation="tanh", seed=0,
                        fused_impl="tcgen05", precision="auto", ctc_weight=0.3, e_lambda=10.0, backward_mode="recompute"):
    """One training step of the hot path on `device` (joint -> RNNT loss, CTC head -> CTC loss, mixed loss
    backward, EWC penalty sweep) and the same step through the CPU oracle.  Returns error metrics."""
    from indic_cl_asr_b200 import CTCLoss, ConvASRDecoder, RNNTJoint, RNNTLoss
    from indic_cl_asr_b200 import cl
    from oracle import cl_oracle, ctc_oracle, joint_oracle

    torch.manual_seed(seed)
    joint = RNNTJoint(jointnet=dict(encoder_hidden=De, pred_hidden=Dp, joint_hidden=H, activation=activation,
                                    dropout=0.0),
                      num_classes=V, fuse_loss_wer=True, fused_batch_size=4, fused_impl=fused_impl,
                      precision=precision, backward_mode=backward_mode).to(device)
    joint.set_loss(RNNTLoss(num_classes=V))
    joint.set_wer(object())
    head = ConvASRDecoder(feat_in=De, num_classes=V).to(device)
    ctc = CTCLoss(num_classes=V, zero_infinity=True)
    model = torch.nn.ModuleDict({"joint": joint, "ctc_decoder": head})

    enc, dec, tr, el, tl = synth_batch(B, T, U, V, De, Dp, seed=seed + 1)
    enc_d = enc.to(device).requires_grad_(True)
    dec_d = dec.to(device).requires_grad_(True)
    tr_d, el_d, tl_d = tr.to(device), el.to(device), tl.to(device)

    # EWC state from a previous "task"
    theta = cl.get_params(model)
    star = cl.get_params_clone(model)
    star.flat.add_(0.01 * torch.randn_like(star.flat))
    fish = cl.get_zero_params(model, device)
    fish.flat.uniform_(0.0, 1.0)

    loss_rnnt, _, _, _ = joint(encoder_outputs=enc_d, decoder_outputs=dec_d, encoder_lengths=el_d,
                               transcripts=tr_d, transcript_lengths=tl_d, compute_wer=False)
    log_probs = head(encoder_output=enc_d)
    loss_ctc = ctc(log_probs=log_probs, targets=tr_d, input_lengths=el_d, target_lengths=tl_d)
    loss = (1 - ctc_weight) * loss_rnnt + ctc_weight * loss_ctc
    pen, pen_avg = cl.get_penalty_grads({"cl_config": {"e_lambda": e_lambda}}, fish, theta, star)
    cl.set_grads(model, pen)
    loss.backward()
    torch.cuda.synchronize()
    got = {n: p.grad.detach().cpu().numpy().copy() for n, p in model.named_parameters()}
    got_loss = float(loss.item())

    # ---------------- oracle (float64 on CPU)
    p64 = {k: v.detach().cpu().double().requires_grad_(True) for k, v in joint_params(joint).items()}
    cw = head.decoder_layers[0].weight.detach().cpu().double().requires_grad_(True)
    cb = head.decoder_layers[0].bias.detach().cpu().double().requires_grad_(True)
    enc64 = enc.double().requires_grad_(True)
    dec64 = dec.double().requires_grad_(True)
    o_rnnt = joint_oracle.fused_joint_loss(enc64, dec64, el, tr, tl, p64, activation, V, 4, "mean_batch")
    lp, _ = joint_oracle.ctc_head(enc64, cw, cb)

    class _Ctc(torch.autograd.Function):
        @staticmethod
        def forward(ctx, lp_):
            nll, g = ctc_oracle.ctc_loss_and_grad(lp_.detach().numpy(), tr.numpy(), el.numpy(), tl.numpy(), V, True)
            ctx.g = torch.from_numpy(g)
            return torch.from_numpy(nll)

        @staticmethod
        def backward(ctx, go):
            return ctx.g * go.view(-1, 1, 1)

    o_ctc = _Ctc.apply(lp).mean()
    o_loss = (1 - ctc_weight) * o_rnnt + ctc_weight * o_ctc
    o_loss.backward()
    names = {"joint.enc.weight": p64["enc.weight"], "joint.enc.bias": p64["enc.bias"],
             "joint.pred.weight": p64["pred.weight"], "joint.pred.bias": p64["pred.bias"],
             f"joint.joint_net.{len(joint.joint_net) - 1}.weight": p64["out.weight"],
             f"joint.joint_net.{len(joint.joint_net) - 1}.bias": p64["out.bias"],
             "ctc_decoder.decoder_layers.0.weight": cw, "ctc_decoder.decoder_layers.0.bias": cb}
    th_c = {k: v.detach().cpu() for k, v in theta.items()}
    st_c = {k: v.detach().cpu() for k, v in star.items()}
    fi_c = {k: v.detach().cpu() for k, v in fish.items()}
    o_pen, o_avg = cl_oracle.get_penalty_grads(e_lambda, fi_c, th_c, st_c)
    worst, worst_name = 0.0, ""
    for n, t in names.items():
        ref = t.grad.numpy() + o_pen[n].double().numpy()
        e = rel_err(got[n], ref)
        if e > worst:
            worst, worst_name = e, n
    return {
        "loss": got_loss, "oracle_loss": float(o_loss.item()),
        "loss_rel_err": abs(got_loss - float(o_loss.item())) / abs(float(o_loss.item())),
        "grad_rel_err": worst, "worst_param": worst_name,
        "d_enc_rel_err": rel_err(enc_d.grad.cpu().numpy(), enc64.grad.numpy()),
        "d_dec_rel_err": rel_err(dec_d.grad.cpu().numpy(), dec64.grad.numpy()),
        "penalty_avg_rel_err": abs(pen_avg - o_avg) / max(abs(o_avg), 1e-30),
    }
