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


@pytest.mark.parametrize("B,K,T", [(3, 20, 7), (4, 512, 250), (2, 33, 65), (5, 640, 101)])
@pytest.mark.parametrize("precision", ["bf16x3", "fp16x3"])
def test_linear_on_nemo_layout_view(B, K, T, precision):
    """The joint and the CTC head receive [B, D, T] tensors and transpose them (reference modules/rnnt.py:1457-1459,
    conv_asr.py:467): linear_x3 recognises the transposed view and goes through clasr_transpose_last2 in both directions.
    The result and the gradient w.r.t. the [B, D, T] leaf (contiguous, like the leaf) must match the plain path bit for bit
    in the forward values and to summation order in the weight gradient."""
    from indic_cl_asr_b200 import linear as L

    N = 48
    g = torch.Generator().manual_seed(B * K + T)
    base = torch.randn(B, K, T, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    dy = torch.randn(B, T, N, generator=g)

    def run(force_copy):
        xb = base.to(DEV).requires_grad_(True)
        wg, bg = w.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
        view = xb.transpose(1, 2)
        assert L._is_transposed_view(view)
        y = linear_x3(view.contiguous() if force_copy else view, wg, bg, precision)
        y.backward(dy.to(DEV))
        torch.cuda.synchronize()
        return y.detach().cpu(), xb.grad, wg.grad.cpu(), bg.grad.cpu()

    y1, gx1, gw1, gb1 = run(False)
    y0, gx0, gw0, gb0 = run(True)
    assert gx1.is_contiguous() and gx1.shape == (B, K, T)
    assert torch.equal(y1, y0)
    assert torch.equal(gx1.cpu(), gx0.cpu())
    assert _rel(gw1, gw0.double()) <= 1e-6 and _rel(gb1, gb0.double()) <= 1e-6
    # and against fp64
    xd, wd, bd = base.double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
    ref = torch.nn.functional.linear(xd.transpose(1, 2), wd, bd)
    ref.backward(dy.double())
    assert _rel(y1, ref.detach()) <= 2e-5
    assert _rel(gx1, xd.grad) <= 2e-5
