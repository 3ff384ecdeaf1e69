"""HybridRNNTCTCLoss / ewc_backward / mas_importance_backward (the training_step loss half and the drivers' inner-loop
steps, reference hybrid_rnnt_ctc_models.py:868-902, cl_baseline_ewc.py:228-255, cl_baseline_mas.py:257-271) against
the step-by-step composition of the same modules and the oracle-checked helpers."""
import pytest
import torch

from helpers import rel_err, synth_batch
from indic_cl_asr_b200 import CTCLoss, ConvASRDecoder, HybridRNNTCTCLoss, RNNTJoint, RNNTLoss, cl
from indic_cl_asr_b200.hybrid import ewc_backward, mas_importance_backward

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _build(V=40, H=64, De=32, Dp=32, act="tanh"):
    torch.manual_seed(0)
    joint = RNNTJoint(jointnet=dict(encoder_hidden=De, pred_hidden=Dp, joint_hidden=H, activation=act, dropout=0.0),
                      num_classes=V, fuse_loss_wer=True, fused_batch_size=4).to(DEV)
    joint.set_loss(RNNTLoss(num_classes=V))
    joint.set_wer(object())
    head = ConvASRDecoder(feat_in=De, num_classes=V).to(DEV)
    ctc = CTCLoss(num_classes=V, zero_infinity=True)
    return joint, head, ctc, torch.nn.ModuleDict({"joint": joint, "ctc_decoder": head})


@pytest.mark.parametrize("overlap", [True, False])
def test_hybrid_loss_equals_manual_composition(overlap):
    joint, head, ctc, model = _build()
    enc, dec, tr, el, tl = synth_batch(6, 25, 8, 40, 32, 32, seed=3, device=DEV)
    step = HybridRNNTCTCLoss(joint, head, ctc, ctc_loss_weight=0.3, overlap_ctc=overlap)
    e1, d1 = enc.clone().requires_grad_(True), dec.clone().requires_grad_(True)
    loss, mon = step(e1, el, d1, tr, tl)
    loss.backward()
    torch.cuda.synchronize()
    got = {n: p.grad.clone() for n, p in model.named_parameters()}
    for p in model.parameters():
        p.grad = None
    e2, d2 = enc.clone().requires_grad_(True), dec.clone().requires_grad_(True)
    l_r, _, _, _ = joint(encoder_outputs=e2, decoder_outputs=d2, encoder_lengths=el, transcripts=tr,
                         transcript_lengths=tl, compute_wer=False)
    l_c = ctc(log_probs=head(encoder_output=e2), targets=tr, input_lengths=el, target_lengths=tl)
    ref = 0.7 * l_r + 0.3 * l_c
    ref.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - ref.item()) <= 1e-6 * abs(ref.item())
    assert torch.allclose(mon["train_rnnt_loss"], l_r.detach()) and torch.allclose(mon["train_ctc_loss"], l_c.detach())
    assert all(v is None or v.is_cuda for v in mon.values())            # no host sync inside the step
    for n, p in model.named_parameters():
        assert rel_err(got[n].cpu().numpy(), p.grad.cpu().numpy()) <= 2e-5, n   # fp32 atomics (split-K, shared reductions) reorder low bits
    assert rel_err(e1.grad.cpu().numpy(), e2.grad.cpu().numpy()) <= 2e-5


def test_ewc_backward_preloads_penalty():
    joint, head, ctc, model = _build()
    enc, dec, tr, el, tl = synth_batch(4, 18, 6, 40, 32, 32, seed=5, device=DEV)
    step = HybridRNNTCTCLoss(joint, head, ctc)
    star = cl.get_params_clone(model)
    star.flat.add_(0.01 * torch.randn_like(star.flat))
    fish = cl.get_zero_params(model, DEV)
    fish.flat.uniform_(0.0, 1.0)
    cfg = {"cl_config": {"e_lambda": 10.0}}
    loss, _ = step(enc.clone().requires_grad_(True), el, dec.clone().requires_grad_(True), tr, tl)
    avg = ewc_backward(model, loss, cfg, fish, star)
    torch.cuda.synchronize()
    got = {n: p.grad.clone() for n, p in model.named_parameters()}
    # reference sequence (cl_baseline_ewc.py:228-240) with the dict API
    for p in model.parameters():
        p.grad = None
    loss2, _ = step(enc.clone().requires_grad_(True), el, dec.clone().requires_grad_(True), tr, tl)
    pen, pen_avg = cl.get_penalty_grads(cfg, fish, cl.get_params(model), star)
    cl.set_grads(model, pen)
    loss2.backward()
    torch.cuda.synchronize()
    assert abs(avg.item() - pen_avg) <= 1e-6 * abs(pen_avg)
    for n, p in model.named_parameters():
        assert rel_err(got[n].cpu().numpy(), p.grad.cpu().numpy()) <= 2e-5, n


def test_mas_importance_backward_accumulates_abs_grad():
    joint, head, ctc, model = _build(act="relu")
    enc, dec, tr, el, tl = synth_batch(5, 16, 5, 40, 32, 32, seed=9, device=DEV)
    step = HybridRNNTCTCLoss(joint, head, ctc)
    joint.store_sub_enc, joint.store_sub_logits, joint.detach_sub_enc = False, True, False
    head.return_logits_ = True
    imp = cl.get_zero_params(model, DEV)
    step(enc.clone().requires_grad_(True), el, dec.clone().requires_grad_(True), tr, tl)
    obj = mas_importance_backward(model, joint, head, imp, mas_ctx=0.3)
    torch.cuda.synchronize()
    assert torch.isfinite(obj) and obj.item() > 0
    for n, p in model.named_parameters():
        assert torch.equal(imp[n], p.grad.abs()), n
