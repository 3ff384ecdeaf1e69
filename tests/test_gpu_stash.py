"""The two modes of the fused joint's backward pass must agree: dZ swept from the logits the forward call kept
(`stash`, include/clasr_b200.h: clasr_joint_stash_bytes) vs dZ from the tile-wise recompute (stash == NULL).
Both read bit-identical logits, so the only differences are fp32 summation orders (bias gradient, split-K)."""
import pytest
import torch

from helpers import rel_err
from indic_cl_asr_b200 import _lib
from indic_cl_asr_b200.fused import fused_joint_rnnt_loss, fused_joint_sumsq
from test_gpu_fused import make

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _grads(fn, tensors):
    leaves = [x.to(DEV).requires_grad_(True) for x in tensors]
    out = fn(*leaves)
    out.backward()
    torch.cuda.synchronize()
    return [x.grad.cpu().numpy() for x in leaves], out.detach().cpu().numpy()


@pytest.mark.parametrize("B,T,U,V,H,act,kw", [
    (3, 37, 11, 300, 128, "tanh", {}),
    (2, 33, 12, 1024, 640, "tanh", {"fastemit_lambda": 0.01}),
    (4, 21, 9, 29, 64, "sigmoid", {"clamp": 0.05}),
    (2, 50, 20, 1024, 320, "relu", {"dropout_p": 0.2, "dropout_seed": 1234}),
    (2, 19, 7, 4500, 64, "tanh", {}),          # more than 512 eight-column groups: two column chunks in the sweep
])
@pytest.mark.parametrize("precision", ["bf16x3", "bf16", "fp16x3", "fp16m8"])
def test_loss_backward_modes_agree(B, T, U, V, H, act, kw, precision, monkeypatch):
    f, g, W, b, lab, al, ll = make(B, T, U, V, H, seed=3 * B + T + V)
    wts = torch.linspace(0.5, 1.5, B).to(DEV)

    def run(fd, gd, Wd, bd):
        c = fused_joint_rnnt_loss(fd, gd, Wd, bd, lab.to(DEV), al.to(DEV), ll.to(DEV), V, act, precision, **kw)
        return (c * wts).sum()

    monkeypatch.setenv("CLASR_JOINT_STASH", "48")
    g_stash, c_stash = _grads(run, (f, g, W, b))
    monkeypatch.setenv("CLASR_JOINT_STASH", "0")
    g_rec, c_rec = _grads(run, (f, g, W, b))
    assert c_stash == c_rec   # the forward statistics do not depend on the mode
    for name, a, r in zip(["d_f", "d_g", "d_W", "d_b"], g_stash, g_rec):
        assert rel_err(a, r) <= 2e-5, name


@pytest.mark.parametrize("B,T,U,V,H,act", [(2, 13, 5, 40, 64, "tanh"), (3, 30, 9, 1024, 640, "tanh")])
def test_sumsq_backward_modes_agree(B, T, U, V, H, act, monkeypatch):
    f, g, W, b, lab, al, ll = make(B, T, U, V, H, seed=B + T)
    up = torch.rand(B, T, U + 1).to(DEV)

    def run(fd, gd, Wd, bd):
        s = fused_joint_sumsq(fd, gd, Wd, bd, lab.to(DEV), al.to(DEV), ll.to(DEV), V, act, "bf16x3")
        return (s * up).sum()

    monkeypatch.setenv("CLASR_JOINT_STASH", "48")
    g_stash, v_stash = _grads(run, (f, g, W, b))
    monkeypatch.setenv("CLASR_JOINT_STASH", "0")
    g_rec, v_rec = _grads(run, (f, g, W, b))
    assert v_stash == v_rec
    for name, a, r in zip(["d_f", "d_g", "d_W", "d_b"], g_stash, g_rec):
        assert rel_err(a, r) <= 2e-5, name


def test_stash_limit_falls_back_to_recompute(monkeypatch):
    """A stash larger than CLASR_JOINT_STASH GiB is not allocated: the call still succeeds (recompute mode)."""
    from indic_cl_asr_b200 import fused

    B, T, U, V, H = 2, 40, 10, 256, 64
    need = _lib.lib().clasr_joint_stash_bytes(B, T, U + 1, H, V + 1, _lib.PREC["bf16x3"])
    assert need > 0
    monkeypatch.setenv("CLASR_JOINT_STASH", str((need - 1) / (1 << 30)))
    assert fused._stash(torch.empty(1, device=DEV), B, T, U + 1, H, V + 1, _lib.PREC["bf16x3"], True) == (None, 0)
    f, g, W, b, lab, al, ll = make(B, T, U, V, H, seed=5)
    fd = f.to(DEV).requires_grad_(True)
    c = fused_joint_rnnt_loss(fd, g.to(DEV), W.to(DEV), b.to(DEV), lab.to(DEV), al.to(DEV), ll.to(DEV), V)
    c.sum().backward()
    torch.cuda.synchronize()
    assert torch.isfinite(fd.grad).all()
