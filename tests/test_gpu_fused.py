"""Fused tcgen05 joint + transducer loss vs the fp64 oracle."""
import numpy as np
import pytest
import torch

from helpers import rel_err
from indic_cl_asr_b200.fused import fused_joint_forward_stats, fused_joint_rnnt_loss
from oracle import joint_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

LOSS_TOL = {"bf16x3": 1e-5, "bf16": 2e-3}
GRAD_TOL = {"bf16x3": 1e-4, "bf16": 3e-2}


def make(B, T, U, V, H, seed, ragged=True):
    g = torch.Generator().manual_seed(seed)
    f = torch.randn(B, T, H, generator=g)
    gg = torch.randn(B, U + 1, H, generator=g)
    W = (torch.rand(V + 1, H, generator=g) * 2 - 1) / H ** 0.5
    b = (torch.rand(V + 1, generator=g) * 2 - 1) / H ** 0.5
    lab = torch.randint(0, V, (B, U), generator=g)
    if ragged:
        al = torch.randint(max(1, T // 2), T + 1, (B,), generator=g); al[0] = T
        ll = torch.randint(0, U + 1, (B,), generator=g); ll[0] = U
    else:
        al, ll = torch.full((B,), T), torch.full((B,), U)
    return f, gg, W, b, lab, al, ll


def oracle(f, g, W, b, lab, al, ll, V, act):
    f64, g64 = f.double().requires_grad_(True), g.double().requires_grad_(True)
    W64, b64 = W.double().requires_grad_(True), b.double().requires_grad_(True)
    h = joint_oracle._ACTS[act](f64.unsqueeze(2) + g64.unsqueeze(1))
    z = torch.nn.functional.linear(h, W64, b64)
    costs = joint_oracle.rnnt_loss(z, lab, al, ll, V)
    return costs, z, (f64, g64, W64, b64)


@pytest.mark.parametrize("precision", ["bf16", "bf16x3"])
@pytest.mark.parametrize("B,T,U,V,H,act", [(2, 9, 4, 20, 64, "tanh"), (3, 40, 17, 256, 128, "relu"),
                                           (2, 33, 12, 1024, 640, "tanh"), (4, 21, 9, 300, 320, "sigmoid")])
def test_forward_costs_and_sumsq(B, T, U, V, H, act, precision):
    f, g, W, b, lab, al, ll = make(B, T, U, V, H, seed=B + T + V)
    costs, ssq = fused_joint_forward_stats(f.to(DEV), g.to(DEV), W.to(DEV), b.to(DEV), lab.to(DEV), al.to(DEV),
                                           ll.to(DEV), V, act, precision)
    torch.cuda.synchronize()
    oc, z, _ = oracle(f, g, W, b, lab, al, ll, V, act)
    assert rel_err(costs.cpu().numpy(), oc.detach().numpy()) <= LOSS_TOL[precision]
    ref_ssq = (z.detach() ** 2).sum(-1).numpy()
    got = ssq.cpu().numpy()
    for i in range(B):
        Tb, Ub1 = int(al[i]), int(ll[i]) + 1
        assert rel_err(got[i, :Tb, :Ub1], ref_ssq[i, :Tb, :Ub1]) <= (1e-5 if precision == "bf16x3" else 2e-2)


@pytest.mark.parametrize("precision", ["bf16", "bf16x3"])
@pytest.mark.parametrize("B,T,U,V,H,act", [(2, 9, 4, 20, 64, "tanh"), (3, 40, 17, 256, 128, "relu"),
                                           (2, 33, 12, 1024, 640, "tanh")])
def test_backward(B, T, U, V, H, act, precision):
    f, g, W, b, lab, al, ll = make(B, T, U, V, H, seed=7 * B + T + V)
    fd, gd, Wd, bd = [x.to(DEV).requires_grad_(True) for x in (f, g, W, b)]
    costs = fused_joint_rnnt_loss(fd, gd, Wd, bd, lab.to(DEV), al.to(DEV), ll.to(DEV), V, act, precision)
    wts = torch.linspace(0.5, 1.5, B)
    (costs * wts.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    oc, _, leaves = oracle(f, g, W, b, lab, al, ll, V, act)
    (oc * wts.double()).sum().backward()
    for name, got, ref in zip(["d_f", "d_g", "d_W", "d_b"], (fd, gd, Wd, bd), leaves):
        assert rel_err(got.grad.cpu().numpy(), ref.grad.numpy()) <= GRAD_TOL[precision], name
