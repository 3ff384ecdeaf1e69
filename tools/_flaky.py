import os, sys, torch
sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo")
from helpers import rel_err
from indic_cl_asr_b200.fused import fused_joint_rnnt_loss
DEV = "cuda:0"
H = 640
B, T_, U_, V_, act = 4, 400, 80, 256, sys.argv[1] if len(sys.argv) > 1 else "relu"
g = torch.Generator().manual_seed(17)
f = (torch.randn(B, T_, H, generator=g) * 0.7).to(DEV)
gg = (torch.randn(B, U_ + 1, H, generator=g) * 0.7).to(DEV)
W = ((torch.rand(V_ + 1, H, generator=g) * 2 - 1) / H ** 0.5).to(DEV)
b = ((torch.rand(V_ + 1, generator=g) * 2 - 1) / H ** 0.5).to(DEV)
lab = torch.randint(0, V_, (B, U_), generator=g).to(DEV)
al = torch.randint(T_ // 2, T_ + 1, (B,), generator=g); al[0] = T_
ll = torch.randint(U_ // 2, U_ + 1, (B,), generator=g); ll[0] = U_
al, ll = al.to(DEV), ll.to(DEV)
ref = None
for it in range(int(os.environ.get("N", "30"))):
    leaves = [x.clone().requires_grad_(True) for x in (f, gg, W, b)]
    costs = fused_joint_rnnt_loss(*leaves, lab, al, ll, V_, act, "bf16x3")
    costs.sum().backward()
    torch.cuda.synchronize()
    cur = [x.grad.clone() for x in leaves] + [costs.detach().clone()]
    if ref is None:
        ref = cur
    else:
        errs = [rel_err(c.cpu().numpy(), r.cpu().numpy()) for c, r in zip(cur, ref)]
        if max(errs) > 1e-5:
            bad = (cur[0] - ref[0]).abs()
            idx = (bad > 1e-5 * ref[0].abs().max()).nonzero()
            print(it, ["%.2e" % e for e in errs], "d_f bad elems", idx.shape[0], idx[:3].tolist(), idx[-1:].tolist())
print("done", act, os.environ.get("CLASR_JOINT_PAIR"))
