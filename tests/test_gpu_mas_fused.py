"""MAS importance pass on the fused path: store_sub_logits -> LazySubLogits answering the driver's objective
(reference cl_baseline_mas.py:212-221, 257-271) from the kernel's per-cell sum_v z^2, vs an fp64 CPU restatement
of the same objective over materialised sub-batch logits (reference modules/rnnt.py:1425-1496, 1617-1650)."""
import pytest
import torch

from helpers import rel_err
from indic_cl_asr_b200 import ConvASRDecoder, RNNTJoint, RNNTLoss
from indic_cl_asr_b200.fused import LazySubLogits

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CTX = 0.3  # config.yaml mas_ctx


def _oracle(params, enc, dec, enc_lens, tr_lens, fbs, activation):
    """fp64: per sub-batch materialised logits (narrowed to the sub-batch maxima), driver objective, autograd."""
    p = {k: v.detach().cpu().double().requires_grad_(True) for k, v in params.items()}
    enc = enc.detach().cpu().double().transpose(1, 2)
    dec = dec.detach().cpu().double().transpose(1, 2)
    act = {"tanh": torch.tanh, "relu": torch.relu, "sigmoid": torch.sigmoid}[activation]
    f = enc @ p["enc.weight"].t() + p["enc.bias"]
    g = dec @ p["pred.weight"].t() + p["pred.bias"]
    B = enc.shape[0]
    terms = []
    for b0 in range(0, B, fbs):
        b1 = min(B, b0 + fbs)
        mt, mu = int(enc_lens[b0:b1].max()), int(tr_lens[b0:b1].max())
        h = act(f[b0:b1, :mt].unsqueeze(2) + g[b0:b1, : mu + 1].unsqueeze(1))
        z = h @ p["out.weight"].t() + p["out.bias"]
        terms.append((z.flatten(end_dim=-2) ** 2).sum(dim=-1).mean())
    rnn = sum(terms) / len(terms)
    zc = enc @ p["ctc.weight"].squeeze(-1).t() + p["ctc.bias"]
    ctc = (zc.flatten(end_dim=-2) ** 2).sum(dim=-1).mean()
    loss = rnn * (1 - CTX) + ctc * CTX
    loss.backward()
    return loss.item(), {k: v.grad for k, v in p.items()}


@pytest.mark.parametrize("activation", ["tanh", "relu"])
@pytest.mark.parametrize("pair", ["1", "0"])
def test_mas_importance_objective_fused_vs_oracle(activation, pair, monkeypatch):
    monkeypatch.setenv("CLASR_JOINT_PAIR", pair)
    torch.manual_seed(3)
    B, T, U, V, H, De, Dp, fbs = 7, 23, 9, 50, 128, 48, 40, 3
    joint = RNNTJoint(jointnet=dict(encoder_hidden=De, pred_hidden=Dp, joint_hidden=H, activation=activation,
                                    dropout=0.0), num_classes=V, fuse_loss_wer=True, fused_batch_size=fbs).to(DEV)
    joint.set_loss(RNNTLoss(num_classes=V))
    joint.set_wer(object())
    head = ConvASRDecoder(feat_in=De, num_classes=V).to(DEV)
    enc = torch.randn(B, De, T, device=DEV)
    dec = torch.randn(B, Dp, U + 1, device=DEV)
    tr = torch.randint(0, V, (B, U), device=DEV)
    enc_lens = torch.tensor([23, 11, 17, 20, 23, 9, 14], device=DEV)
    tr_lens = torch.tensor([9, 4, 7, 8, 2, 5, 9], device=DEV)

    # the driver's importance epoch (cl_baseline_mas.py:212-221, 257-265)
    joint.store_sub_enc, joint.store_sub_logits, joint.detach_sub_enc = False, True, False
    head.return_logits_ = True
    loss_rnnt, _, _, _ = joint(encoder_outputs=enc, decoder_outputs=dec, encoder_lengths=enc_lens, transcripts=tr,
                               transcript_lengths=tr_lens, compute_wer=False)
    head(encoder_output=enc)
    assert torch.isfinite(loss_rnnt)
    assert all(isinstance(x, LazySubLogits) for x in joint.store_list)        # logits were never materialised
    assert [tuple(x.shape) for x in joint.store_list] == [(3, 23, 10, V + 1), (3, 23, 9, V + 1), (1, 14, 10, V + 1)]
    decoder_logits = (head.decoder_logits.flatten(end_dim=-2) ** 2).sum(dim=-1).mean()
    rnn_logits = 0
    for i in joint.store_list:
        rnn_logits += (i.flatten(end_dim=-2) ** 2).sum(dim=-1).mean()
    rnn_logits /= len(joint.store_list)
    loss = rnn_logits * (1 - CTX) + decoder_logits * CTX
    loss.backward()

    params = {"enc.weight": joint.enc.weight, "enc.bias": joint.enc.bias, "pred.weight": joint.pred.weight,
              "pred.bias": joint.pred.bias, "out.weight": joint.joint_net[-1].weight, "out.bias": joint.joint_net[-1].bias,
              "ctc.weight": head.decoder_layers[0].weight, "ctc.bias": head.decoder_layers[0].bias}
    ref_loss, ref_grads = _oracle(params, enc, dec, enc_lens.cpu(), tr_lens.cpu(), fbs, activation)
    assert abs(loss.item() - ref_loss) <= 1e-5 * abs(ref_loss)
    for k, p in params.items():
        assert rel_err(p.grad.cpu().double().numpy(), ref_grads[k].numpy()) <= 1e-4, k


def test_lazy_sub_logits_materialise_fallback():
    """Anything other than the sum-of-squares chain pays for the logits and matches the materialising joint."""
    torch.manual_seed(5)
    B, T, U, V, H, De, Dp = 2, 6, 3, 10, 64, 16, 16
    joint = RNNTJoint(jointnet=dict(encoder_hidden=De, pred_hidden=Dp, joint_hidden=H, activation="tanh", dropout=0.0),
                      num_classes=V, fuse_loss_wer=True, fused_batch_size=2).to(DEV)
    joint.set_loss(RNNTLoss(num_classes=V))
    joint.set_wer(object())
    enc, dec = torch.randn(B, De, T, device=DEV), torch.randn(B, Dp, U + 1, device=DEV)
    joint.store_sub_logits, joint.detach_sub_enc = True, False
    joint(encoder_outputs=enc, decoder_outputs=dec, encoder_lengths=torch.tensor([6, 6], device=DEV),
          transcripts=torch.randint(0, V, (B, U), device=DEV), transcript_lengths=torch.tensor([3, 3], device=DEV),
          compute_wer=False)
    lazy = joint.store_list[0]
    z = lazy.materialise()
    assert z.shape == lazy.shape == (2, 6, 4, V + 1)
    assert torch.allclose((lazy ** 2).sum(dim=-1), (z ** 2).sum(dim=-1), rtol=1e-4, atol=1e-5)
    assert torch.allclose(lazy.abs().max(), z.abs().max())  # __getattr__ fallback
