"""BASELINE.json full size (T=250, U=100, V=1024, H=640): the fused tcgen05 path in its default precision against
the fp64 oracle (two utterances), against the materialised path at B=32 (torch fp32 joint -> our RNNT loss kernels,
itself pinned to the fp64 oracle at this size by test_gpu_rnnt_loss.py::test_full_size_properties) and against
size-independent properties of the transducer gradient."""
import pytest
import torch

from helpers import rel_err
from indic_cl_asr_b200 import RNNTLossNumba
from indic_cl_asr_b200.fused import fused_joint_rnnt_loss

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
T, U, V, H = 250, 100, 1024, 640


def _inputs(B, seed, ragged):
    g = torch.Generator().manual_seed(seed)
    f = torch.randn(B, T, H, generator=g) * 0.7
    gg = torch.randn(B, U + 1, H, generator=g) * 0.7
    W = (torch.rand(V + 1, H, generator=g) * 2 - 1) / H ** 0.5
    b = (torch.rand(V + 1, generator=g) * 2 - 1) / H ** 0.5
    lab = torch.randint(0, V, (B, U), generator=g)
    if ragged:
        al = torch.randint(T // 2, T + 1, (B,), generator=g); al[0] = T
        ll = torch.randint(U // 2, U + 1, (B,), generator=g); ll[0] = U
    else:
        al, ll = torch.full((B,), T), torch.full((B,), U)
    return [x.to(DEV) for x in (f, gg, W, b, lab, al, ll)]


def _materialised_costs(f, g, W, b, lab, al, ll, sub=4):
    """torch fp32 joint in sub-batches of 4 (like the reference's fused_batch_size) + our drop-in RNNT loss."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        loss = RNNTLossNumba(blank=V, reduction="none")
        out = []
        for b0 in range(0, f.shape[0], sub):
            sl = slice(b0, b0 + sub)
            mt, mu = int(al[sl].max()), int(ll[sl].max())
            z = torch.nn.functional.linear(torch.tanh(f[sl, :mt].unsqueeze(2) + g[sl, : mu + 1].unsqueeze(1)), W, b)
            out.append(loss(z, lab[sl, :mu].contiguous(), al[sl], ll[sl]))
        return torch.cat(out)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


# "fp16m8" is what RNNTJoint(precision="auto") resolves to, i.e. what bench.py measures
@pytest.mark.parametrize("precision", ["fp16m8", "fp16x3", "bf16x3"])
@pytest.mark.parametrize("ragged", [False, True])
def test_full_size_costs_fused_vs_materialised(ragged, precision):
    f, g, W, b, lab, al, ll = _inputs(32, 11, ragged)
    with torch.no_grad():
        fused = fused_joint_rnnt_loss(f, g, W, b, lab, al, ll, V, "tanh", precision)
        ref = _materialised_costs(f, g, W, b, lab, al, ll)
    assert torch.isfinite(fused).all()
    assert ((fused - ref).abs() <= 1e-5 * ref.abs()).all(), ((fused - ref).abs() / ref.abs()).max().item()


@pytest.mark.parametrize("precision,stash_gib", [("fp16m8", None), ("fp16m8", 48.0), ("fp16x3", None), ("bf16x3", None)])
def test_full_size_gradients_and_properties(precision, stash_gib):
    B = 8
    f, g, W, b, lab, al, ll = _inputs(B, 5, True)
    wts = torch.linspace(0.5, 1.5, B, device=DEV)
    leaves = [x.clone().requires_grad_(True) for x in (f, g, W, b)]
    (fused_joint_rnnt_loss(*leaves, lab, al, ll, V, "tanh", precision, stash_gib=stash_gib) * wts).sum().backward()
    ref_leaves = [x.clone().requires_grad_(True) for x in (f, g, W, b)]
    (_materialised_costs(*ref_leaves, lab, al, ll) * wts).sum().backward()
    for name, got, ref in zip(["d_f", "d_g", "d_W", "d_b"], leaves, ref_leaves):
        assert rel_err(got.grad.cpu().numpy(), ref.grad.cpu().numpy()) <= 1e-4, name
    d_f, d_g, d_W, d_b = [x.grad for x in leaves]
    # flow conservation: every row of dZ sums to zero over the vocabulary => sum_v d_b[v] = 0
    assert abs(d_b.double().sum().item()) <= 1e-4 * d_b.abs().max().item() * 32
    # padding never contributes: frames beyond T_b and prediction rows beyond U_b get exactly-zero gradient
    for i in range(B):
        assert d_f[i, int(al[i]):].abs().sum().item() == 0.0
        assert d_g[i, int(ll[i]) + 1:].abs().sum().item() == 0.0
    # blank mass: each alignment emits exactly T_b blanks => sum over cells of dZ[., blank] = sum_b w_b (E[#blank] - T_b)
    # with E[#blank] = sum of blank posteriors, which the softmax part reproduces; only sign/finite checks here
    assert torch.isfinite(d_W).all() and d_b[V].item() < 0.0


@pytest.mark.parametrize("precision", ["fp16x3", "fp16m8"])
@pytest.mark.parametrize("stash_gib", [None, 48.0], ids=["recompute", "stash"])
@pytest.mark.parametrize("act", ["tanh", "relu"])
def test_config2_default_precision_vs_fp64_oracle(act, stash_gib, precision):
    """BASELINE.json configs[1] shape (T=250, U=100, V=1024, H=640), the DEFAULT precision (fp16m8) and fp16x3, both backward
    modes, tanh and the shipped checkpoint's ReLU — against the fp64 ORACLE itself (oracle/joint_oracle.py +
    rnnt_oracle.py), not against this repo's materialised path.  Two utterances (one full length, one ragged) keep the
    fp64 joint (2 x 250 x 101 x 1025 logits) and the numpy lattice at a few seconds of CPU.
    Tolerances are north_star's: relative 1e-5 on the loss, 1e-4 on every gradient tensor."""
    from oracle import joint_oracle

    g_ = torch.Generator().manual_seed(23)
    B = 2
    f = torch.randn(B, T, H, generator=g_) * 0.7
    gg = torch.randn(B, U + 1, H, generator=g_) * 0.7
    W = (torch.rand(V + 1, H, generator=g_) * 2 - 1) / H ** 0.5
    b = (torch.rand(V + 1, generator=g_) * 2 - 1) / H ** 0.5
    lab = torch.randint(0, V, (B, U), generator=g_)
    al, ll = torch.tensor([T, 173]), torch.tensor([U, 61])
    wts = torch.tensor([0.35, 0.65])          # mean_batch x loss weight: upstream gradients well below 1
    leaves = [x.to(DEV).requires_grad_(True) for x in (f, gg, W, b)]
    costs = fused_joint_rnnt_loss(*leaves, lab.to(DEV), al.to(DEV), ll.to(DEV), V, act, precision, stash_gib=stash_gib)
    (costs * wts.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    ref_leaves = [x.double().requires_grad_(True) for x in (f, gg, W, b)]
    z = torch.nn.functional.linear(joint_oracle._ACTS[act](ref_leaves[0].unsqueeze(2) + ref_leaves[1].unsqueeze(1)),
                                   ref_leaves[2], ref_leaves[3])
    oc = joint_oracle.rnnt_loss(z, lab, al, ll, V)
    (oc * wts.double()).sum().backward()
    errs = {nm: rel_err(got.grad.cpu().numpy(), rf.grad.numpy()) for nm, got, rf in zip(["d_f", "d_g", "d_W", "d_b"], leaves, ref_leaves)}
    errs["cost"] = rel_err(costs.detach().cpu().numpy(), oc.detach().numpy())
    print(f"config2 {act} {precision} stash={stash_gib}: " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()))
    assert errs["cost"] <= 1e-5
    for nm in ("d_f", "d_g", "d_W", "d_b"):
        assert errs[nm] <= 1e-4, (act, nm, errs)


def _generic_costs(f, g, W, b, lab, al, ll, V, act, sub=2):
    loss = RNNTLossNumba(blank=V, reduction="none")
    fn = {"tanh": torch.tanh, "relu": torch.relu}[act]
    out = []
    for b0 in range(0, f.shape[0], sub):
        sl = slice(b0, b0 + sub)
        mt, mu = int(al[sl].max()), int(ll[sl].max())
        z = torch.nn.functional.linear(fn(f[sl, :mt].unsqueeze(2) + g[sl, : mu + 1].unsqueeze(1)), W, b)
        out.append(loss(z, lab[sl, :mu].contiguous(), al[sl], ll[sl]))
    return torch.cat(out)


@pytest.mark.parametrize("name,B,T_,U_,V_,act,precision,ltol,gtol", [
    # configs[2]: IndicConformer-medium shapes (16 s audio -> T'~400, per-language vocabulary 256, ReLU joint)
    ("config3", 4, 400, 80, 256, "relu", "bf16x3", 1e-5, 1e-4),
    ("config3_default_precision", 4, 400, 80, 256, "relu", "fp16m8", 1e-5, 1e-4),
    # configs[4]: V=4096 multilingual tokenizer, T=500, U=200, bf16 joint GEMM (single MMA term: bf16 tolerances)
    ("config5", 2, 500, 200, 4096, "tanh", "bf16", 2e-3, 3e-2),
])
def test_other_config_shapes_fused_vs_materialised(name, B, T_, U_, V_, act, precision, ltol, gtol):
    g = torch.Generator().manual_seed(17)
    f = (torch.randn(B, T_, H, generator=g) * 0.7).to(DEV)
    gg = (torch.randn(B, U_ + 1, H, generator=g) * 0.7).to(DEV)
    W = ((torch.rand(V_ + 1, H, generator=g) * 2 - 1) / H ** 0.5).to(DEV)
    b = ((torch.rand(V_ + 1, generator=g) * 2 - 1) / H ** 0.5).to(DEV)
    lab = torch.randint(0, V_, (B, U_), generator=g).to(DEV)
    al = torch.randint(T_ // 2, T_ + 1, (B,), generator=g); al[0] = T_
    ll = torch.randint(U_ // 2, U_ + 1, (B,), generator=g); ll[0] = U_
    al, ll = al.to(DEV), ll.to(DEV)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        leaves = [x.clone().requires_grad_(True) for x in (f, gg, W, b)]
        costs = fused_joint_rnnt_loss(*leaves, lab, al, ll, V_, act, precision)
        costs.sum().backward()
        ref_leaves = [x.clone().requires_grad_(True) for x in (f, gg, W, b)]
        ref = _generic_costs(*ref_leaves, lab, al, ll, V_, act)
        ref.sum().backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    assert ((costs - ref).abs() <= ltol * ref.abs()).all(), ((costs - ref).abs() / ref.abs()).max().item()
    for nm, got, rf in zip(["d_f", "d_g", "d_W", "d_b"], leaves, ref_leaves):
        assert rel_err(got.grad.cpu().numpy(), rf.grad.cpu().numpy()) <= gtol, (name, nm)
