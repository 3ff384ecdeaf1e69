"""In-kernel joint dropout (reference modules/rnnt.py:1699-1709: act -> Dropout(p) -> Linear) on the fused path.

torch's Philox stream cannot be reproduced bit-for-bit, so parity is checked with the SAME mask: the kernels' counter-
based mask is re-derived on the host (fused.dropout_mask_reference) and applied inside the fp64 oracle."""
import numpy as np
import pytest
import torch

from helpers import rel_err
from indic_cl_asr_b200 import RNNTJoint, RNNTLoss
from indic_cl_asr_b200.fused import dropout_mask_reference, fused_joint_rnnt_loss
from oracle import joint_oracle
from test_gpu_fused import make

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("pair", ["1", "0"])
@pytest.mark.parametrize("B,T,U,V,H,act,p", [(3, 21, 8, 60, 128, "tanh", 0.2), (2, 40, 17, 300, 640, "relu", 0.35),
                                             (4, 9, 3, 20, 64, "sigmoid", 0.1)])
def test_fused_dropout_matches_oracle_with_same_mask(B, T, U, V, H, act, p, pair, monkeypatch):
    monkeypatch.setenv("CLASR_JOINT_PAIR", pair)
    f, g, W, b, lab, al, ll = make(B, T, U, V, H, seed=3 * B + T)
    seed = 0x1234_5678_9ABC_DEF1 + T
    fd, gd, Wd, bd = [x.to(DEV).requires_grad_(True) for x in (f, g, W, b)]
    costs = fused_joint_rnnt_loss(fd, gd, Wd, bd, lab.to(DEV), al.to(DEV), ll.to(DEV), V, act, "bf16x3",
                                  dropout_p=p, dropout_seed=seed)
    wts = torch.linspace(0.5, 1.5, B)
    (costs * wts.to(DEV)).sum().backward()
    torch.cuda.synchronize()

    keep, scale = dropout_mask_reference(al, ll, T, U + 1, H, p, seed)
    # the mask is a fair coin at rate p over the valid cells
    valid = np.zeros((B, T, U + 1), dtype=bool)
    for i in range(B):
        valid[i, : int(al[i]), : int(ll[i]) + 1] = True
    rate = 1.0 - keep[valid].mean()
    assert abs(rate - p) < 4 * (p * (1 - p) / keep[valid].size) ** 0.5 + 1e-3, rate

    f64, g64 = f.double().requires_grad_(True), g.double().requires_grad_(True)
    W64, b64 = W.double().requires_grad_(True), b.double().requires_grad_(True)
    h = joint_oracle._ACTS[act](f64.unsqueeze(2) + g64.unsqueeze(1)) * torch.from_numpy(keep).double() * scale
    oc = joint_oracle.rnnt_loss(torch.nn.functional.linear(h, W64, b64), lab, al, ll, V)
    (oc * wts.double()).sum().backward()
    assert rel_err(costs.detach().cpu().numpy(), oc.detach().numpy()) <= 1e-5
    for name, got, ref in zip(["d_f", "d_g", "d_W", "d_b"], (fd, gd, Wd, bd), (f64, g64, W64, b64)):
        assert rel_err(got.grad.cpu().numpy(), ref.grad.numpy()) <= 1e-4, name


def test_joint_module_uses_fused_dropout_in_training():
    """RNNTJoint(dropout=0.2).train() stays on the tcgen05 path (the shipped checkpoint's setting); eval() == p=0."""
    torch.manual_seed(0)
    V, H = 30, 64
    joint = RNNTJoint(jointnet=dict(encoder_hidden=32, pred_hidden=32, joint_hidden=H, activation="relu", dropout=0.2),
                      num_classes=V, fuse_loss_wer=True, fused_batch_size=4).to(DEV)
    joint.set_loss(RNNTLoss(num_classes=V))
    joint.set_wer(object())
    assert joint._tcgen05_supported(None)
    enc, dec = torch.randn(4, 32, 15, device=DEV), torch.randn(4, 32, 6, device=DEV)
    tr = torch.randint(0, V, (4, 5), device=DEV)
    el, tl = torch.tensor([15, 12, 9, 15], device=DEV), torch.tensor([5, 3, 4, 2], device=DEV)
    kw = dict(encoder_outputs=enc, decoder_outputs=dec, encoder_lengths=el, transcripts=tr, transcript_lengths=tl,
              compute_wer=False)
    joint.train()
    torch.manual_seed(7); a = joint(**kw)[0]
    torch.manual_seed(7); b = joint(**kw)[0]
    torch.manual_seed(8); c = joint(**kw)[0]
    assert torch.equal(a, b) and not torch.equal(a, c)        # reproducible under torch.manual_seed, seed-dependent
    joint.eval()
    e1, e2 = joint(**kw)[0], joint(**kw)[0]
    assert torch.equal(e1, e2) and not torch.equal(e1, a)
    joint.train()
    loss = joint(**kw)[0]
    loss.backward()
    assert all(torch.isfinite(p.grad).all() for p in joint.parameters())
