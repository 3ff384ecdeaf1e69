"""GPU parity of RNNTJoint (both fused strategies) and the full step against the reference-run fixtures / oracle."""
import numpy as np
import pytest
import torch

from conftest import split_cases
from helpers import rel_err, run_step_and_oracle
from indic_cl_asr_b200 import ConvASRDecoder, RNNTJoint, RNNTLoss

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def build_from_fixture(c, fused_impl, precision="auto", backward_mode="recompute"):
    """RNNTJoint + RNNTLoss configured and loaded like the reference run that produced fixture case ``c``
    (oracle/gen_golden.py::_joint_case); ``precision='auto'`` is the module default that bench.py runs."""
    B, T, U, De, Dp, H, V, fbs = [int(x) for x in c["cfg"]]
    multi = "language_keys" in c
    keys = [str(k) for k in c["language_keys"]] if multi else None
    j = RNNTJoint(jointnet=dict(encoder_hidden=De, pred_hidden=Dp, joint_hidden=H, activation=str(c["activation"]),
                                dropout=0.0),
                  num_classes=V * len(keys) if multi else V, fuse_loss_wer=True, fused_batch_size=fbs,
                  fused_impl=fused_impl, precision=precision, backward_mode=backward_mode, multilingual=multi,
                  language_keys=keys)
    sd = {"enc.weight": c["p.enc.weight"], "enc.bias": c["p.enc.bias"], "pred.weight": c["p.pred.weight"],
          "pred.bias": c["p.pred.bias"]}
    if multi:
        for k in keys:
            sd[f"joint_net.1.{k}.weight"] = c[f"p.out.{k}.weight"]
            sd[f"joint_net.1.{k}.bias"] = c[f"p.out.{k}.bias"]
    else:
        sd.update({"joint_net.1.weight": c["p.out.weight"], "joint_net.1.bias": c["p.out.bias"]})
    j.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})
    j = j.to(DEV)
    kw = {}
    if "fastemit_lambda" in c and (float(c["fastemit_lambda"]) != 0.0 or float(c["clamp"]) > 0.0):
        kw = {"loss_kwargs": dict(fastemit_lambda=float(c["fastemit_lambda"]), clamp=float(c["clamp"]))}
    j.set_loss(RNNTLoss(num_classes=V, reduction=str(c["reduction"]), **kw))
    j.set_wer(object())
    return j


def fixture_param_grads(j, c):
    """{fixture key: parameter} for every parameter the fixture holds a gradient for."""
    got = {"enc.weight": j.enc.weight, "enc.bias": j.enc.bias, "pred.weight": j.pred.weight, "pred.bias": j.pred.bias}
    if "language_keys" in c:
        for k in [str(x) for x in c["language_keys"]]:
            got[f"out.{k}.weight"] = j.joint_net[1][k].weight
            got[f"out.{k}.bias"] = j.joint_net[1][k].bias
    else:
        got.update({"out.weight": j.joint_net[1].weight, "out.bias": j.joint_net[1].bias})
    return got


def run_fixture(j, c):
    enc = torch.tensor(c["enc"], device=DEV, requires_grad=True)
    dec = torch.tensor(c["dec"], device=DEV, requires_grad=True)
    kw = {"language_ids": [str(x) for x in c["language_ids"]]} if "language_ids" in c else {}
    loss, _, _, _ = j(encoder_outputs=enc, decoder_outputs=dec, encoder_lengths=torch.tensor(c["enc_lens"], device=DEV),
                      transcripts=torch.tensor(c["transcripts"], device=DEV),
                      transcript_lengths=torch.tensor(c["transcript_lens"], device=DEV), compute_wer=False, **kw)
    loss.backward()
    torch.cuda.synchronize()
    return loss, enc, dec


def assert_fixture_parity(j, c, loss, enc, dec):
    assert abs(loss.item() - float(c["loss"])) <= 1e-5 * abs(float(c["loss"]))          # north_star: rel 1e-5 on loss
    for k, p in fixture_param_grads(j, c).items():
        ref = c["g." + k]
        got = p.grad.cpu().numpy() if p.grad is not None else np.zeros_like(ref)  # a head no utterance used: grad None
        assert np.abs(got - ref).max() <= 1e-4 * max(np.abs(ref).max(), 1e-6), k         # rel 1e-4 on gradients
    assert rel_err(enc.grad.cpu().numpy(), c["d_enc"]) <= 1e-4
    assert rel_err(dec.grad.cpu().numpy(), c["d_dec"]) <= 1e-4


FUSED_CASES = ["relu_h64", "relu_h128_fastemit", "tanh_h128_mean_volume", "sigmoid_h64_mean", "tanh_h64_fastemit_sum",
               "multilingual_relu", "multilingual_mixed"]


@pytest.mark.parametrize("case", FUSED_CASES)
@pytest.mark.parametrize("impl,precision,mode", [("tcgen05", "auto", "recompute"), ("tcgen05", "auto", "stash"),
                                                 ("tcgen05", "bf16x3", "recompute"), ("materialised", "auto", "recompute")])
def test_fused_shapes_vs_reference_run(golden, case, impl, precision, mode):
    """ref_joint_fused.npz — the reference's own RNNTJoint + RNNTLoss run at joint_hidden in {64, 128}: ReLU (the
    shipped checkpoint), the multilingual per-language head with language_ids (reference modules/rnnt.py:1627-1639,
    1694-1703), FastEmit, mean / mean_volume / sum reductions — against the fused tcgen05 path in its default
    precision (what bench.py runs) and both backward modes, and against the materialised strategy."""
    c = split_cases(golden("ref_joint_fused.npz"))[case]
    j = build_from_fixture(c, impl, precision, mode)
    if impl == "tcgen05" and case != "multilingual_mixed":
        lang = [str(x) for x in c["language_ids"]] if "language_ids" in c else None
        assert j._tcgen05_supported(lang), "this fixture is meant to reach the fused kernel"
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)  # multilingual_mixed: documented fallback to the sub-batch loop
        loss, enc, dec = run_fixture(j, c)
    assert_fixture_parity(j, c, loss, enc, dec)


def test_mixed_language_batch_warns_and_falls_back(golden):
    c = split_cases(golden("ref_joint_fused.npz"))["multilingual_mixed"]
    j = build_from_fixture(c, "tcgen05")
    with pytest.warns(RuntimeWarning, match="falls back to the materialised strategy"):
        run_fixture(j, c)


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


@pytest.mark.parametrize("precision", ["auto", "bf16x3"])
@pytest.mark.parametrize("activation,H", [("tanh", 64), ("relu", 128), ("sigmoid", 64)])
def test_full_step_tcgen05_vs_oracle(activation, H, precision):
    """The B200 path: fused tcgen05 joint + wavefront + CTC + EWC sweep, one step, against the fp64 oracle."""
    rep = run_step_and_oracle(device=DEV, B=5, T=23, U=9, V=37, H=H, De=24, Dp=16, activation=activation, seed=5,
                              fused_impl="tcgen05", precision=precision)
    assert rep["loss_rel_err"] <= 1e-5, rep
    assert rep["grad_rel_err"] <= 1e-4, rep
    assert rep["d_enc_rel_err"] <= 1e-4 and rep["d_dec_rel_err"] <= 1e-4, rep


@pytest.mark.parametrize("precision", ["auto", "bf16x3"])
def test_tcgen05_matches_reference_run_fixture(golden, precision):
    """ref_joint.npz 'tanh_wide' (H=64): the reference's own RNNTJoint+RNNTLoss output vs the fused B200 path."""
    c = split_cases(golden("ref_joint.npz"))["tanh_wide"]
    j = build_from_fixture(c, "tcgen05", precision)
    loss, enc, dec = run_fixture(j, c)
    assert_fixture_parity(j, c, loss, enc, dec)


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
