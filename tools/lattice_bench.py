"""Times the alpha/beta wavefront (and the CTC lattice) alone: materialised RNNT loss at a tiny vocabulary, so that the
lattice kernels dominate.  Developer tool (run under gpurun, optionally under ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from indic_cl_asr_b200 import _lib
from indic_cl_asr_b200.losses.ctc import ctc_loss
from indic_cl_asr_b200.losses.rnnt import rnnt_loss

B, T, U, V = [int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (32, 250, 100, 15))]
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
z = torch.randn(B, T, U + 1, V + 1, generator=g).to(dev)
lab = torch.randint(0, V, (B, U), generator=g).to(dev)
al = torch.full((B,), T, dtype=torch.long, device=dev)
ll = torch.full((B,), U, dtype=torch.long, device=dev)
lp = torch.randn(B, T, 1025, generator=g).log_softmax(-1).to(dev).requires_grad_(True)
lab_c = torch.randint(0, 1024, (B, U), generator=g).to(dev)
L = _lib.lib()
for sel in ("generic", "shfl"):
    os.environ["CLASR_LATTICE"] = sel
    os.environ["CLASR_CTC_LATTICE"] = sel
    for _ in range(3):
        c = rnnt_loss(z, lab, al, ll, V, None, 0.0, 0.0)
        n = ctc_loss(lp, lab_c, al, ll, 1024, True)
    torch.cuda.synchronize()
    L.clasr_set_profiling(1)
    L.clasr_profile_reset()
    for _ in range(10):
        c = rnnt_loss(z, lab, al, ll, V, None, 0.0, 0.0)
        n = ctc_loss(lp, lab_c, al, ll, 1024, True)
    torch.cuda.synchronize()
    print(sel, "rnnt_lattice ms", _lib.profile_ms("rnnt_lattice"), "ctc_lattice ms", _lib.profile_ms("ctc_lattice"),
          "cost0", float(c[0]), "nll0", float(n[0]))
    L.clasr_set_profiling(0)
