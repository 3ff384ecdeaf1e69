"""Run-to-run repeatability of the fused joint forward/backward.

The kernels are warp-specialised pipelines over shared memory / tensor memory; a missing hand-shake shows up as an
occasional corrupted row, not as a systematic error, so a single comparison against the oracle can pass by luck.
(A row table shared by two producer warps did exactly that: one wrong dZ row in ~1 of 3 runs at config-3 shapes.)
Here the same inputs are pushed through 20 times: costs must be bit-identical, gradients equal up to the
reordering of fp32 atomics (split-K, shared-memory reductions)."""
import pytest
import torch

from helpers import rel_err
from indic_cl_asr_b200.fused import fused_joint_rnnt_loss

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("pair", ["1", "0"])
@pytest.mark.parametrize("B,T,U,V,H,act", [(4, 400, 80, 256, 640, "relu"),     # 3 N tiles: producers barely gated
                                           (6, 250, 100, 1024, 640, "tanh"),   # headline shape
                                           (5, 37, 11, 50, 128, "sigmoid")])   # tiny tiles, many null/partial tiles
def test_fused_joint_is_repeatable(B, T, U, V, H, act, pair, monkeypatch):
    monkeypatch.setenv("CLASR_JOINT_PAIR", pair)
    g = torch.Generator().manual_seed(17)
    f = (torch.randn(B, T, H, generator=g) * 0.7).to(DEV)
    gg = (torch.randn(B, U + 1, H, generator=g) * 0.7).to(DEV)
    W = ((torch.rand(V + 1, H, generator=g) * 2 - 1) / H ** 0.5).to(DEV)
    b = ((torch.rand(V + 1, generator=g) * 2 - 1) / H ** 0.5).to(DEV)
    lab = torch.randint(0, V, (B, U), generator=g).to(DEV)
    al = torch.randint(T // 2, T + 1, (B,), generator=g); al[0] = T
    ll = torch.randint(U // 2, U + 1, (B,), generator=g); ll[0] = U
    al, ll = al.to(DEV), ll.to(DEV)
    ref = None
    for it in range(20):
        leaves = [x.clone().requires_grad_(True) for x in (f, gg, W, b)]
        costs = fused_joint_rnnt_loss(*leaves, lab, al, ll, V, act, "bf16x3")
        costs.sum().backward()
        torch.cuda.synchronize()
        cur = [x.grad.clone() for x in leaves]
        if ref is None:
            ref, ref_costs = cur, costs.detach().clone()
            continue
        assert torch.equal(costs.detach(), ref_costs), it
        for name, c, r in zip(["d_f", "d_g", "d_W", "d_b"], cur, ref):
            assert rel_err(c.cpu().numpy(), r.cpu().numpy()) <= 2e-5, (it, name)
