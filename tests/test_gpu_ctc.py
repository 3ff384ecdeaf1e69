"""GPU parity of the CTC kernels against the reference's goldens, the oracle and torch.nn.CTCLoss."""
import numpy as np
import pytest
import torch

from conftest import split_cases
from helpers import rel_err
from indic_cl_asr_b200 import CTCLoss
from indic_cl_asr_b200.losses.ctc import ctc_loss
from indic_cl_asr_b200.modules.conv_asr import log_softmax_rows
from oracle import ctc_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("name", ["small", "blank_last"])
def test_kat(golden, name):
    k = golden("ref_kat.npz")
    logits = torch.tensor(k[f"ctc_{name}_acts"], dtype=torch.float32, device=DEV, requires_grad=True)
    labels = torch.tensor(k[f"ctc_{name}_labels"], device=DEV)
    V = logits.shape[-1]
    blank = 0 if name == "small" else V - 1
    lp = log_softmax_rows(logits)
    nll = ctc_loss(lp, labels, torch.tensor([logits.shape[1]], device=DEV), torch.tensor([labels.shape[1]], device=DEV),
                   blank, True)
    nll.sum().backward()
    assert np.allclose(nll.sum().item(), k[f"ctc_{name}_expected_cost"], rtol=1e-6)
    assert np.allclose(logits.grad.cpu().numpy(), k[f"ctc_{name}_expected_grads"], atol=1e-6)


@pytest.mark.parametrize("case", ["mean_batch", "mean_volume", "infeasible", "repeats"])
def test_vs_reference_run(golden, case):
    c = split_cases(golden("ref_ctc.npz"))[case]
    V = int(c["num_classes"])
    logits = torch.tensor(c["logits"], device=DEV, requires_grad=True)
    lp = log_softmax_rows(logits)
    lp.retain_grad()
    loss = CTCLoss(num_classes=V, zero_infinity=True, reduction=str(c["reduction"]))(
        log_probs=lp, targets=torch.tensor(c["targets"], device=DEV),
        input_lengths=torch.tensor(c["input_lens"], device=DEV), target_lengths=torch.tensor(c["target_lens"], device=DEV))
    loss.backward()
    assert np.allclose(loss.item(), c["loss"], rtol=1e-5)
    assert np.allclose(lp.grad.cpu().numpy(), c["d_log_probs"], atol=2e-6)
    assert np.allclose(logits.grad.cpu().numpy(), c["d_logits"], atol=2e-6)


@pytest.mark.parametrize("B,T,U,V", [(4, 50, 20, 30), (2, 300, 100, 1024), (3, 8, 0, 5), (2, 40, 19, 3)])
def test_random_vs_oracle_and_torch(B, T, U, V):
    g = torch.Generator().manual_seed(B * 100 + T)
    logits = 2.0 * torch.randn(B, T, V + 1, generator=g)
    targets = torch.randint(0, V, (B, max(U, 1)), generator=g)[:, :U].reshape(B, U)
    il = torch.randint(max(1, T // 2), T + 1, (B,), generator=g); il[0] = T
    tl = torch.randint(0, U + 1, (B,), generator=g); tl[0] = U
    x = logits.to(DEV).requires_grad_(True)
    lp = log_softmax_rows(x)
    lp.retain_grad()
    nll = ctc_loss(lp, targets.to(DEV), il.to(DEV), tl.to(DEV), V, True)
    nll.sum().backward()
    lp64 = torch.log_softmax(logits.double(), -1)
    on, og = ctc_oracle.ctc_loss_and_grad(lp64.numpy(), targets.numpy(), il.numpy(), tl.numpy(), V, True)
    assert np.allclose(nll.detach().cpu().numpy(), on, rtol=1e-5, atol=1e-5)
    assert rel_err(lp.grad.cpu().numpy(), og) <= 1e-4
    # torch.nn.CTCLoss itself (the reference's third-party arithmetic), CPU
    xt = logits.clone().requires_grad_(True)
    lpt = torch.log_softmax(xt, -1)
    t_nll = torch.nn.functional.ctc_loss(lpt.transpose(0, 1), targets, il, tl, blank=V, reduction="none",
                                         zero_infinity=True)
    t_nll.sum().backward()
    assert np.allclose(nll.detach().cpu().numpy(), t_nll.detach().numpy(), rtol=1e-5, atol=1e-5)
    # ATen carries alpha/beta in fp32 log space: at T=300, V=1024 (|alpha| ~ 2e3, ulp 2.4e-4) its own gradient is
    # ~1e-3 away from the fp64 truth that the kernels here match to 1e-4 (asserted above); small cases agree to 1e-4.
    assert rel_err(x.grad.cpu().numpy(), xt.grad.numpy()) <= (1e-4 if T * (V + 1) < 100000 else 3e-3)


def test_log_softmax_rows():
    x = torch.randn(37, 1025, device=DEV, requires_grad=True)
    y = log_softmax_rows(x)
    w = torch.randn_like(y)
    (y * w).sum().backward()
    xr = x.detach().clone().requires_grad_(True)
    yr = torch.log_softmax(xr, -1)
    (yr * w).sum().backward()
    assert torch.allclose(y, yr, atol=1e-6) and torch.allclose(x.grad, xr.grad, atol=1e-5)
