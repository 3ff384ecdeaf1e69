"""Tiny fused joint forward/backward (+ MAS sum-of-squares backward) for compute-sanitizer runs:
    compute-sanitizer --tool racecheck python tools/sanitize_case.py
    compute-sanitizer --tool memcheck  python tools/sanitize_case.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from indic_cl_asr_b200.fused import fused_joint_rnnt_loss, fused_joint_sumsq

DEV = "cuda:0"
for pair in ("1", "0"):
    os.environ["CLASR_JOINT_PAIR"] = pair
    for (B, T, U, V, H, act) in [(3, 19, 6, 40, 64, "tanh"), (2, 33, 9, 300, 128, "relu")]:
        g = torch.Generator().manual_seed(1)
        f = torch.randn(B, T, H, generator=g).to(DEV).requires_grad_(True)
        gg = torch.randn(B, U + 1, H, generator=g).to(DEV).requires_grad_(True)
        W = (torch.randn(V + 1, H, generator=g) / H ** 0.5).to(DEV).requires_grad_(True)
        b = torch.zeros(V + 1, device=DEV, requires_grad=True)
        lab = torch.randint(0, V, (B, U), generator=g).to(DEV)
        al = torch.tensor([T] + [max(1, T - 3 * i) for i in range(1, B)], device=DEV)
        ll = torch.tensor([U] + [max(0, U - 2 * i) for i in range(1, B)], device=DEV)
        c = fused_joint_rnnt_loss(f, gg, W, b, lab, al, ll, V, act, "bf16x3")
        c.sum().backward()
        s = fused_joint_sumsq(f, gg, W, b, lab, al, ll, V, act, "bf16x3")
        s.mean().backward()
        torch.cuda.synchronize()
        print(pair, B, T, U, V, H, act, float(c.sum()), float(s.mean()))
print("ok")
