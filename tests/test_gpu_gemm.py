"""tcgen05 GEMM building block vs an fp64 reference (torch CPU)."""
import numpy as np
import pytest
import torch

from indic_cl_asr_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def gemm_nt(A, B, precision):
    M, K = A.shape
    N = B.shape[0]
    L = _lib.lib()
    prec = _lib.PREC[precision]
    nbytes = L.clasr_gemm_workspace_bytes(M, N, K, prec)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    C = torch.full((M, N), float("nan"), device=DEV)
    _lib.check(L.clasr_gemm_nt(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, prec, ws.data_ptr(), nbytes,
                               _lib.stream_ptr()), "gemm_nt")
    torch.cuda.synchronize()
    return C


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 640), (300, 520, 200), (1000, 1025, 640), (37, 19, 13),
                                   (4096, 640, 1040)])
@pytest.mark.parametrize("precision,tol", [("bf16", 2e-2), ("bf16x3", 2e-5), ("fp16x3", 5e-6)])
def test_gemm_nt(M, N, K, precision, tol):
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g) / K ** 0.5
    C = gemm_nt(A.to(DEV), B.to(DEV), precision).cpu().double()
    ref = A.double() @ B.double().t()
    assert torch.isfinite(C).all()
    err = (C - ref).abs().max().item() / ref.abs().max().item()
    assert err <= tol, err


def gemm_ex(A, B, M, N, K, a_trans, b_trans, k_splits, precision):
    L = _lib.lib()
    prec = _lib.PREC[precision]
    nbytes = L.clasr_gemm_workspace_bytes(M, N, K, prec)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    C = torch.full((M, N), float("nan"), device=DEV)
    _lib.check(L.clasr_gemm_ex(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, int(a_trans), int(b_trans), k_splits,
                               prec, ws.data_ptr(), nbytes, _lib.stream_ptr()), "gemm_ex")
    torch.cuda.synchronize()
    return C


@pytest.mark.parametrize("a_trans,b_trans", [(0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K,splits", [(128, 256, 64, 1), (200, 300, 500, 1), (1025, 648, 4000, 7), (640, 1025, 256, 1)])
@pytest.mark.parametrize("precision,tol", [("bf16", 2e-2), ("bf16x3", 2e-5), ("fp16x3", 5e-6)])
def test_gemm_transposed_operands_and_split_k(M, N, K, splits, a_trans, b_trans, precision, tol):
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g) / K ** 0.5
    Ain = (A.t().contiguous() if a_trans else A).to(DEV)
    Bin = (B.t().contiguous() if b_trans else B).to(DEV)
    C = gemm_ex(Ain, Bin, M, N, K, a_trans, b_trans, splits, precision).cpu().double()
    ref = A.double() @ B.double().t()
    assert torch.isfinite(C).all()
    err = (C - ref).abs().max().item() / ref.abs().max().item()
    assert err <= tol, err


def _to_m8_range(x):
    """fp16m8 operands must sit at max|x| in [2^13, 2^14) (include/clasr_b200.h): exact power-of-two scale."""
    e = int(np.ceil(np.log2(float(x.abs().max()))))
    if 2.0 ** e == float(x.abs().max()):
        e += 1
    return x * 2.0 ** (14 - e), 2.0 ** (14 - e)


@pytest.mark.parametrize("a_trans,b_trans", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K,splits", [(256, 256, 64, 1), (512, 640, 1088, 1), (1024, 640, 4096, 7), (300, 200, 500, 1),
                                          (128, 256, 64, 1)])
def test_gemm_fp16m8(M, N, K, splits, a_trans, b_trans):
    """fp16 hi.hi + two dense e4m3 correction MMAs (kind::f8f6f4) in the CTA-pair GEMM, every operand orientation the
    joint backward uses (dHid: A K-major, B MN-major; dW: both MN-major) plus the plain NT form: ~1e-5 of max|C|
    (tools/fp8_const_scale_study.py), 30x better than a single fp16 pass."""
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g) / K ** 0.5
    A, sa = _to_m8_range(A)
    B, sb = _to_m8_range(B)
    Ain = (A.t().contiguous() if a_trans else A).to(DEV)
    Bin = (B.t().contiguous() if b_trans else B).to(DEV)
    C = gemm_ex(Ain, Bin, M, N, K, a_trans, b_trans, splits, "fp16m8").cpu().double()
    ref = A.double() @ B.double().t()
    assert torch.isfinite(C).all()
    err = (C - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 3e-5, err
