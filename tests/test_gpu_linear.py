"""linear_x3 (tcgen05 Linear: enc / pred projections, CTC-head k=1 conv) vs an fp64 torch reference."""
import pytest
import torch

from indic_cl_asr_b200.linear import linear_x3

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    return (a.double().cpu() - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


@pytest.mark.parametrize("lead,K,N", [((4, 25), 32, 64), ((32, 250), 512, 640), ((32, 101), 640, 640),
                                      ((8, 100), 512, 1025), ((3, 7), 20, 9)])
@pytest.mark.parametrize("bias", [True, False])
@pytest.mark.parametrize("precision,fwd_tol", [("bf16x3", 2e-5), ("fp16x3", 3e-6)])   # fp16 halves: the fp32 accumulation floor
def test_linear_fwd_bwd(lead, K, N, bias, precision, fwd_tol):
    g = torch.Generator().manual_seed(K + N)
    x = torch.randn(*lead, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g) if bias else None
    dy = torch.randn(*lead, N, generator=g)
    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    bd = b.double().requires_grad_(True) if bias else None
    ref = torch.nn.functional.linear(xd, wd, bd)
    ref.backward(dy.double())

    xg, wg = x.to(DEV).requires_grad_(True), w.to(DEV).requires_grad_(True)
    bg = b.to(DEV).requires_grad_(True) if bias else None
    # "fp16x3" = fp16 halves in the forward GEMM (fp32-grade pre-activations for ReLU joints), bf16 split in the
    # backward GEMMs with tiny upstream gradients (1e-7: would underflow fp16)
    scale = 1e-7 if precision == "fp16x3" else 1.0
    y = linear_x3(xg, wg, bg, precision)
    y.backward(dy.to(DEV) * scale)
    torch.cuda.synchronize()
    if scale != 1.0:
        xg.grad /= scale
        wg.grad /= scale
        if bias:
            bg.grad /= scale
    assert y.shape == ref.shape
    assert _rel(y, ref.detach()) <= fwd_tol
    assert _rel(xg.grad, xd.grad) <= 2e-5
    assert _rel(wg.grad, wd.grad) <= 2e-5
    if bias:
        assert _rel(bg.grad, bd.grad) <= 2e-5


def test_linear_noncontiguous_input_and_partial_grads():
    """[B,D,T] -> transpose(1,2) (NeMo layout) and a weight that does not require grad."""
    g = torch.Generator().manual_seed(7)
    enc = torch.randn(3, 48, 17, generator=g)
    w = torch.randn(33, 48, generator=g)
    eg = enc.to(DEV).requires_grad_(True)
    y = linear_x3(eg.transpose(1, 2), w.to(DEV), None)
    y.sum().backward()
    ref_in = enc.double().requires_grad_(True)
    ref = torch.nn.functional.linear(ref_in.transpose(1, 2), w.double())
    ref.sum().backward()
    assert _rel(y, ref.detach()) <= 2e-5
    assert _rel(eg.grad, ref_in.grad) <= 2e-5


def test_linear_rejects_cpu():
    with pytest.raises(RuntimeError):
        linear_x3(torch.randn(4, 8), torch.randn(3, 8))
