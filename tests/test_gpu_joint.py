"""GPU parity of RNNTJoint (both fused strategies) and the full step against the reference-run fixtures / oracle."""
import numpy as np
import pytest
import torch

from conftest import split_cases
from helpers import rel_err, run_step_and_oracle
from indic_cl_asr_b200 import ConvASRDecoder, RNNTJoint, RNNTLoss

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def build_from_fixture(c, fused_impl, precision="bf16x3"):
    B, T, U, De, Dp, H, V, fbs = [int(x) for x in c["cfg"]]
    j = RNNTJoint(jointnet=dict(encoder_hidden=De, pred_hidden=Dp, joint_hidden=H, activation=str(c["activation"]),
                                dropout=0.0),
                  num_classes=V, fuse_loss_wer=True, fused_batch_size=fbs, fused_impl=fused_impl, precision=precision)
    sd = {"enc.weight": c["p.enc.weight"], "enc.bias": c["p.enc.bias"], "pred.weight": c["p.pred.weight"],
          "pred.bias": c["p.pred.bias"], "joint_net.1.weight": c["p.out.weight"], "joint_net.1.bias": c["p.out.bias"]}
    j.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})
    j = j.to(DEV)
    j.set_loss(RNNTLoss(num_classes=V, reduction=str(c["reduction"])))
    j.set_wer(object())
    return j


@pytest.mark.parametrize("case", ["tanh", "relu", "sigmoid", "tanh_wide"])
def test_materialised_vs_reference_run(golden, case):
    c = split_cases(golden("ref_joint.npz"))[case]
    j = build_from_fixture(c, "materialised")
    enc = torch.tensor(c["enc"], device=DEV, requires_grad=True)
    dec = torch.tensor(c["dec"], device=DEV, requires_grad=True)
    # non-fused forward returns raw logits on CUDA (modules/rnnt.py:1651-1655); the CPU reference log-softmaxed them
    j._fuse_loss_wer = False
    z = j(encoder_outputs=enc, decoder_outputs=dec)
    assert np.allclose(torch.log_softmax(z, -1).detach().cpu().numpy(), c["logits"], atol=1e-5)
    j._fuse_loss_wer = True
    j.store_sub_logits = True
    loss, wer, _, _ = j(encoder_outputs=enc, decoder_outputs=dec, encoder_lengths=torch.tensor(c["enc_lens"], device=DEV),
                        transcripts=torch.tensor(c["transcripts"], device=DEV),
                        transcript_lengths=torch.tensor(c["transcript_lens"], device=DEV), compute_wer=False)
    assert wer is None
    assert [list(s.shape) for s in j.store_list] == c["sub_shapes"].tolist()
    loss.backward()
    assert abs(loss.item() - float(c["loss"])) <= 1e-5 * abs(float(c["loss"]))
    got = {"enc.weight": j.enc.weight, "enc.bias": j.enc.bias, "pred.weight": j.pred.weight, "pred.bias": j.pred.bias,
           "out.weight": j.joint_net[1].weight, "out.bias": j.joint_net[1].bias}
    for k, p in got.items():
        assert rel_err(p.grad.cpu().numpy(), c["g." + k]) <= 1e-4, k
    assert rel_err(enc.grad.cpu().numpy(), c["d_enc"]) <= 1e-4
    assert rel_err(dec.grad.cpu().numpy(), c["d_dec"]) <= 1e-4


def test_conv_asr_decoder_vs_reference_run(golden):
    c = golden("ref_conv_asr.npz")
    ncls = c["weight"].shape[0]
    d = ConvASRDecoder(feat_in=c["weight"].shape[1], num_classes=ncls - 1,
                       language_masks={"bn": c["mask"].tolist()}).to(DEV)
    d.load_state_dict({"decoder_layers.0.weight": torch.tensor(c["weight"]), "decoder_layers.0.bias": torch.tensor(c["bias"])})
    d.return_logits_ = True
    lp = d(encoder_output=torch.tensor(c["x"], device=DEV), language_ids=["bn", "bn"])
    assert np.allclose(lp.detach().cpu().numpy(), c["log_probs"], atol=1e-5)
    assert np.allclose(d.decoder_logits.detach().cpu().numpy(), c["logits"], atol=1e-5)


@pytest.mark.parametrize("activation", ["tanh", "relu"])
def test_full_step_materialised_vs_oracle(activation):
    rep = run_step_and_oracle(device=DEV, B=5, T=23, U=9, V=37, H=64, De=24, Dp=16, activation=activation, seed=3,
                              fused_impl="materialised")
    assert rep["loss_rel_err"] <= 1e-5, rep
    assert rep["grad_rel_err"] <= 1e-4, rep
    assert rep["d_enc_rel_err"] <= 1e-4 and rep["d_dec_rel_err"] <= 1e-4, rep
    assert rep["penalty_avg_rel_err"] <= 1e-5, rep


@pytest.mark.parametrize("activation,H,precision", [("tanh", 64, "bf16x3"), ("relu", 128, "bf16x3"), ("sigmoid", 64, "bf16x3")])
def test_full_step_tcgen05_vs_oracle(activation, H, precision):
    """The B200 path: fused tcgen05 joint + wavefront + CTC + EWC sweep, one step, against the fp64 oracle."""
    rep = run_step_and_oracle(device=DEV, B=5, T=23, U=9, V=37, H=H, De=24, Dp=16, activation=activation, seed=5,
                              fused_impl="tcgen05", precision=precision)
    assert rep["loss_rel_err"] <= 1e-5, rep
    assert rep["grad_rel_err"] <= 1e-4, rep
    assert rep["d_enc_rel_err"] <= 1e-4 and rep["d_dec_rel_err"] <= 1e-4, rep


def test_tcgen05_matches_reference_run_fixture(golden):
    """ref_joint.npz 'tanh_wide' (H=64): the reference's own RNNTJoint+RNNTLoss output vs the fused B200 path."""
    c = split_cases(golden("ref_joint.npz"))["tanh_wide"]
    j = build_from_fixture(c, "tcgen05")
    enc = torch.tensor(c["enc"], device=DEV, requires_grad=True)
    dec = torch.tensor(c["dec"], device=DEV, requires_grad=True)
    loss, _, _, _ = j(encoder_outputs=enc, decoder_outputs=dec, encoder_lengths=torch.tensor(c["enc_lens"], device=DEV),
                      transcripts=torch.tensor(c["transcripts"], device=DEV),
                      transcript_lengths=torch.tensor(c["transcript_lens"], device=DEV), compute_wer=False)
    loss.backward()
    assert abs(loss.item() - float(c["loss"])) <= 1e-5 * abs(float(c["loss"]))
    got = {"enc.weight": j.enc.weight, "enc.bias": j.enc.bias, "pred.weight": j.pred.weight, "pred.bias": j.pred.bias,
           "out.weight": j.joint_net[1].weight, "out.bias": j.joint_net[1].bias}
    for k, p in got.items():
        assert rel_err(p.grad.cpu().numpy(), c["g." + k]) <= 1e-4, k
    assert rel_err(enc.grad.cpu().numpy(), c["d_enc"]) <= 1e-4
    assert rel_err(dec.grad.cpu().numpy(), c["d_dec"]) <= 1e-4


def test_wer_hook_is_called_per_sub_batch():
    class Wer:
        def __init__(self): self.calls = []
        def update(self, **kw): self.calls.append({k: (v.shape if hasattr(v, "shape") else v) for k, v in kw.items()})
        def compute(self): return torch.tensor(0.5), torch.tensor(1.0), torch.tensor(2.0)
        def reset(self): pass
    j = RNNTJoint(jointnet=dict(encoder_hidden=8, pred_hidden=8, joint_hidden=64, activation="relu"), num_classes=11,
                  fuse_loss_wer=True, fused_batch_size=2).to(DEV)
    w = Wer()
    j.set_loss(RNNTLoss(num_classes=11)); j.set_wer(w)
    B, T, U = 5, 7, 3
    out = j(encoder_outputs=torch.randn(B, 8, T, device=DEV), decoder_outputs=torch.randn(B, 8, U + 1, device=DEV),
            encoder_lengths=torch.tensor([7, 6, 5, 7, 3], device=DEV), transcripts=torch.randint(0, 11, (B, U), device=DEV),
            transcript_lengths=torch.tensor([3, 2, 3, 1, 0], device=DEV), compute_wer=True)
    loss, wer, num, den = out
    assert len(w.calls) == 3 and w.calls[0]["predictions"] == torch.Size([2, 8, 7])  # [B,D,T] like the reference
    assert w.calls[2]["predictions"] == torch.Size([1, 8, 3])
    assert float(wer) == 0.5 and float(num) == 3.0 and float(den) == 6.0 and torch.isfinite(loss)
