"""Fused tcgen05 joint + transducer loss vs the fp64 oracle."""
import numpy as np
import pytest
import torch

from helpers import rel_err
from indic_cl_asr_b200.fused import fused_joint_forward_stats, fused_joint_rnnt_loss
from oracle import joint_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# fp16m8: fp16x3 forward (same costs), backward GEMMs as fp16 + two e4m3 correction terms: north_star's 1e-4 with margin
LOSS_TOL = {"bf16x3": 1e-5, "fp16x3": 1e-5, "bf16": 2e-3, "fp16m8": 1e-5}
GRAD_TOL = {"bf16x3": 1e-4, "fp16x3": 2e-5, "bf16": 3e-2, "fp16m8": 6e-5}


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


@pytest.mark.parametrize("precision", ["bf16", "bf16x3", "fp16x3", "fp16m8"])
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
        assert rel_err(got[i, :Tb, :Ub1], ref_ssq[i, :Tb, :Ub1]) <= (2e-2 if precision == "bf16" else 3e-5 if precision == "fp16m8" else 1e-5)


@pytest.mark.parametrize("precision", ["bf16", "bf16x3", "fp16x3", "fp16m8"])
@pytest.mark.parametrize("B,T,U,V,H,act", [(2, 9, 4, 20, 64, "tanh"), (3, 40, 17, 256, 128, "relu"),
                                           (2, 33, 12, 1024, 640, "tanh")])
@pytest.mark.parametrize("stash", ["48", "0"], ids=["stash", "recompute"])
def test_backward(B, T, U, V, H, act, precision, stash, monkeypatch):
    # both backward modes: dZ from the logits the forward kept (CLASR_JOINT_STASH=<GiB>) / from a tile-wise recompute (the default)
    monkeypatch.setenv("CLASR_JOINT_STASH", stash)
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


@pytest.mark.parametrize("act", ["tanh", "sigmoid", "relu"])
def test_saturated_preactivations(act):
    """|f|, |g| far beyond the range where exp() is finite, with opposite signs meeting: the activation must
    saturate exactly like the fp64 oracle (no clamping of the individual projections)."""
    B, T, U, V, H = 2, 12, 5, 30, 64
    f, g, W, b, lab, al, ll = make(B, T, U, V, H, seed=99, ragged=False)
    f = f * 60.0          # pre-activations up to ~ +-200
    g = g * 60.0
    f[0, 0, :8], g[0, 0, :8] = 150.0, -149.0      # tanh(1) after cancellation of two huge terms
    f[1, 3, 8:16], g[1, 2, 8:16] = -120.0, 121.5
    fd, gd, Wd, bd = [x.to(DEV).requires_grad_(True) for x in (f, g, W, b)]
    costs = fused_joint_rnnt_loss(fd, gd, Wd, bd, lab.to(DEV), al.to(DEV), ll.to(DEV), V, act, "bf16x3")
    costs.sum().backward()
    oc, _, leaves = oracle(f, g, W, b, lab, al, ll, V, act)
    oc.sum().backward()
    assert torch.isfinite(costs).all()
    assert rel_err(costs.detach().cpu().numpy(), oc.detach().numpy()) <= 1e-5
    for name, got, ref in zip(["d_f", "d_g", "d_W", "d_b"], (fd, gd, Wd, bd), leaves):
        assert torch.isfinite(got.grad).all(), name
        assert rel_err(got.grad.cpu().numpy(), ref.grad.numpy()) <= 1e-4, name


def test_degenerate_shapes():
    """Single cell (T=1, U=0), empty transcripts for the whole batch, and T_b = 1 inside a ragged batch."""
    for B, T, U, al, ll in [(1, 1, 0, [1], [0]), (3, 7, 0, [7, 3, 1], [0, 0, 0]), (3, 9, 4, [9, 1, 5], [4, 0, 2])]:
        V, H = 17, 64
        f, g, W, b, lab, _, _ = make(B, T, max(U, 1), V, H, seed=B * 31 + T, ragged=False)
        g = g[:, : U + 1].contiguous()
        lab = lab[:, :U].contiguous()
        al_t, ll_t = torch.tensor(al), torch.tensor(ll)
        fd, gd, Wd, bd = [x.to(DEV).requires_grad_(True) for x in (f, g, W, b)]
        costs = fused_joint_rnnt_loss(fd, gd, Wd, bd, lab.to(DEV), al_t.to(DEV), ll_t.to(DEV), V, "tanh", "bf16x3")
        costs.sum().backward()
        oc, _, leaves = oracle(f, g, W, b, lab, al_t, ll_t, V, "tanh")
        oc.sum().backward()
        assert rel_err(costs.detach().cpu().numpy(), oc.detach().numpy()) <= 1e-5, (B, T, U)
        for name, got, ref in zip(["d_f", "d_g", "d_W", "d_b"], (fd, gd, Wd, bd), leaves):
            assert rel_err(got.grad.cpu().numpy(), ref.grad.numpy()) <= 1e-4, (B, T, U, name)


@pytest.mark.parametrize("scale", [1e-5, 1.0, 3e4])
@pytest.mark.parametrize("stash", ["48", "0"], ids=["stash", "recompute"])
@pytest.mark.parametrize("prec", ["fp16x3", "fp16m8"])
def test_fp16x3_gradient_scale_invariance(scale, stash, prec, monkeypatch):
    """fp16 operands lose relative accuracy below 6e-5, so dZ is produced pre-scaled by a power of two derived from the
    upstream gradient (joint_gscale_kernel) and un-scaled where it is consumed: parity must not depend on the size of
    the upstream gradient (mean reductions / loss weights make it tiny, a loss scale makes it huge)."""
    monkeypatch.setenv("CLASR_JOINT_STASH", stash)
    B, T, U, V, H = 3, 33, 12, 300, 128
    f, g, W, b, lab, al, ll = make(B, T, U, V, H, seed=41)
    fd, gd, Wd, bd = [x.to(DEV).requires_grad_(True) for x in (f, g, W, b)]
    costs = fused_joint_rnnt_loss(fd, gd, Wd, bd, lab.to(DEV), al.to(DEV), ll.to(DEV), V, "tanh", prec,
                                  fastemit_lambda=0.01)
    wts = torch.linspace(0.5, 1.5, B) * scale
    (costs * wts.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    f64, g64 = f.double().requires_grad_(True), g.double().requires_grad_(True)
    W64, b64 = W.double().requires_grad_(True), b.double().requires_grad_(True)
    z = torch.nn.functional.linear(torch.tanh(f64.unsqueeze(2) + g64.unsqueeze(1)), W64, b64)
    oc = joint_oracle.rnnt_loss(z, lab, al, ll, V, fastemit_lambda=0.01)
    (oc * wts.double()).sum().backward()
    for name, got, ref in zip(["d_f", "d_g", "d_W", "d_b"], (fd, gd, Wd, bd), (f64, g64, W64, b64)):
        assert torch.isfinite(got.grad).all(), name
        assert rel_err(got.grad.cpu().numpy(), ref.grad.numpy()) <= GRAD_TOL[prec], (name, scale)


@pytest.mark.parametrize("prec", ["fp16x3", "fp16m8"])
@pytest.mark.parametrize("w_mult", [1e-4, 30.0])
def test_fp16x3_weight_scale_invariance(w_mult, prec):
    """W_out is split into fp16 halves after a power-of-two scale that brings max|W| to ~1 (tiny weights would sit in
    fp16's subnormal range otherwise); logits and gradients are un-scaled where they are consumed."""
    B, T, U, V, H = 2, 21, 9, 300, 128
    f, g, W, b, lab, al, ll = make(B, T, U, V, H, seed=43)
    W = W * w_mult
    fd, gd, Wd, bd = [x.to(DEV).requires_grad_(True) for x in (f, g, W, b)]
    costs = fused_joint_rnnt_loss(fd, gd, Wd, bd, lab.to(DEV), al.to(DEV), ll.to(DEV), V, "tanh", prec)
    costs.sum().backward()
    torch.cuda.synchronize()
    oc, _, leaves = oracle(f, g, W, b, lab, al, ll, V, "tanh")
    oc.sum().backward()
    assert rel_err(costs.detach().cpu().numpy(), oc.detach().numpy()) <= 1e-5
    for name, got, ref in zip(["d_f", "d_g", "d_W", "d_b"], (fd, gd, Wd, bd), leaves):
        assert rel_err(got.grad.cpu().numpy(), ref.grad.numpy()) <= GRAD_TOL[prec], (name, w_mult)


@pytest.mark.parametrize("prec", ["fp16x3", "fp16m8"])
@pytest.mark.parametrize("mult", [1e-3, 300.0])
def test_fp16x3_relu_hidden_scale(mult, prec):
    """ReLU hidden values are unbounded: with fp16 operands the A producers scale them by a power of two derived from
    max|f| + max|g|, and the logits / dW are un-scaled where they are consumed."""
    B, T, U, V, H = 2, 21, 9, 300, 128
    f, g, W, b, lab, al, ll = make(B, T, U, V, H, seed=47)
    f, g = f * mult, g * mult
    W = W / max(mult, 1.0)      # keep the logits O(1) so that the softmax stays informative
    fd, gd, Wd, bd = [x.to(DEV).requires_grad_(True) for x in (f, g, W, b)]
    costs = fused_joint_rnnt_loss(fd, gd, Wd, bd, lab.to(DEV), al.to(DEV), ll.to(DEV), V, "relu", prec)
    costs.sum().backward()
    torch.cuda.synchronize()
    oc, _, leaves = oracle(f, g, W, b, lab, al, ll, V, "relu")
    oc.sum().backward()
    assert rel_err(costs.detach().cpu().numpy(), oc.detach().numpy()) <= 1e-5
    for name, got, ref in zip(["d_f", "d_g", "d_W", "d_b"], (fd, gd, Wd, bd), leaves):
        assert torch.isfinite(got.grad).all(), name
        assert rel_err(got.grad.cpu().numpy(), ref.grad.numpy()) <= GRAD_TOL[prec], (name, mult)
