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
