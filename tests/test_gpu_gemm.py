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
@pytest.mark.parametrize("precision,tol", [("bf16", 2e-2), ("bf16x3", 2e-5)])
def test_gemm_nt(M, N, K, precision, tol):
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g) / K ** 0.5
    C = gemm_nt(A.to(DEV), B.to(DEV), precision).cpu().double()
    ref = A.double() @ B.double().t()
    assert torch.isfinite(C).all()
    err = (C - ref).abs().max().item() / ref.abs().max().item()
    assert err <= tol, err
