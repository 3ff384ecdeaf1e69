"""Cycles per tcgen05.mma on an idle SM (design micro-benchmark; run on a B200)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from indic_cl_asr_b200 import _lib

L = _lib.lib()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
for pattern, name in ((0, "SS"), (1, "TS"), (2, "SS,SS,TS")):
    for N in (48, 64, 96, 128, 160, 192, 256):
        for _ in range(2):
            _lib.check(L.clasr_debug_mma_rate(N, pattern, 2000, out.data_ptr(), _lib.stream_ptr()), "mma_rate")
            torch.cuda.synchronize()
        cyc, n = out.tolist()
        print(f"{name:9s} N={N:3d}: {cyc / n:7.1f} cycles/MMA  (floor N/2 = {N // 2})")
