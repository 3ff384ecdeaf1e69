"""Developer tool: where do the joint kernel's roles wait?  Needs a library built with -DCLASR_TRACE
(see csrc/joint_fused.cu: CLASR_TRACE_WAIT) selected through CLASR_LIB.  Prints, per mode, the mean over CTAs of the
cycles each role spent in its barrier waits as a fraction of the kernel's cycles.

    cd indic_cl_asr_b200/csrc && touch joint_fused.cu && \
      make NVCCFLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC \
                      -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -DCLASR_TRACE" && \
      cp ../libclasr_sm100.so build/libclasr_trace.so && touch joint_fused.cu && make     # product library again
    CLASR_LIB=indic_cl_asr_b200/csrc/build/libclasr_trace.so python tools/joint_trace.py
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from indic_cl_asr_b200 import _lib  # noqa: E402
from indic_cl_asr_b200.fused import fused_joint_rnnt_loss  # noqa: E402

SLOTS = ["MMA<-tmem_empty", "MMA<-a_ready", "MMA<-full", "epi<-tmem_full", "prod<-a_free", "TMA<-empty", "total",
         "epi<-zstore", "prod load+act", "prod lo->TMEM", "prod fence+arrive"]


def read_trace():
    L = _lib.lib()
    if not hasattr(L, "clasr_debug_joint_trace"):
        raise SystemExit("this library was not built with -DCLASR_TRACE: build one (see the module docstring) and "
                         "point CLASR_LIB at it")
    fn = L.clasr_debug_joint_trace
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int]
    buf = (C.c_ulonglong * (148 * 12))()
    rc = fn(buf, 148 * 12)
    assert rc == 0, rc
    return np.array(buf[:], dtype=np.float64).reshape(148, 12)


PRECISION = sys.argv[1] if len(sys.argv) > 1 else "fp16m8"


def main():
    print("precision", PRECISION)
    B, T, U, V, H = 32, 250, 100, 1024, 640
    dev = "cuda:0"
    g = torch.Generator().manual_seed(0)
    f = torch.randn(B, T, H, generator=g).to(dev)
    gg = torch.randn(B, U + 1, H, generator=g).to(dev)
    W = ((torch.rand(V + 1, H, generator=g) * 2 - 1) / H ** 0.5).to(dev)
    b = torch.zeros(V + 1).to(dev)
    lab = torch.randint(0, V, (B, U), generator=g).to(dev)
    al = torch.full((B,), T).to(dev)
    ll = torch.full((B,), U).to(dev)
    for stash, label in (("0", "mode 0 fwd then pass 2a (recompute)"), ("48", "mode 3 fwd (stash)")):
        os.environ["CLASR_JOINT_STASH"] = stash
        for grad in (False, True):
            fd = f.clone().requires_grad_(grad)
            for _ in range(3):
                c = fused_joint_rnnt_loss(fd, gg, W, b, lab, al, ll, V, "tanh", PRECISION)
                if grad:
                    c.sum().backward()
            tr = read_trace()   # last joint_fwd_kernel launch: fwd (no grad) or pass 2a / fwd-stash (grad)
            if stash == "48" and not grad:
                continue
            what = "fwd mode 0" if not grad else ("pass 2a" if stash == "0" else "fwd mode 3")
            lead = tr[0::2]     # leader CTAs own the MMA warp
            tot = tr[:, 6].mean()
            print(f"== {what}: {tot / 1e6:.2f} M cycles")
            for i, n in enumerate(SLOTS):
                if i == 6:
                    continue
                src = lead if i < 3 else tr
                print(f"   {n:18s} {src[:, i].mean() / tot * 100:5.1f} %")


if __name__ == "__main__":
    main()
