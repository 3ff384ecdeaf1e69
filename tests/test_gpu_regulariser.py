"""GPU parity of the EWC / MAS sweeps against the oracle, the reference-run fixture and plain torch ops."""
import numpy as np
import pytest
import torch

from helpers import rel_err
from indic_cl_asr_b200 import cl
from oracle import cl_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def small_model():
    torch.manual_seed(31)
    m = torch.nn.Sequential()
    m.add_module("a", torch.nn.Linear(7, 5))
    m.add_module("frozen", torch.nn.Linear(5, 5))
    m.add_module("b", torch.nn.Linear(5, 3, bias=False))
    for p in m.frozen.parameters():
        p.requires_grad = False
    return m.to(DEV)


def test_reference_fixture(golden):
    c = golden("ref_cl.npz")
    names = [str(n) for n in c["names"]]
    m = small_model()
    theta = cl.get_params(m)
    assert list(theta.keys()) == names  # requires_grad-filtered named_parameters order (utils.py:273)
    for n in names:
        theta[n].copy_(torch.tensor(c["theta." + n]))
    star = cl.get_zero_params(m, DEV); fish = cl.get_zero_params(m, DEV)
    for n in names:
        star[n].copy_(torch.tensor(c["star." + n])); fish[n].copy_(torch.tensor(c["F." + n]))
    pen, avg = cl.get_penalty_grads({"cl_config": {"e_lambda": float(c["e_lambda"])}}, fish, theta, star)
    for n in names:
        assert np.array_equal(pen[n].cpu().numpy(), c["pen." + n]), n   # bit-exact with the reference's torch ops
    assert np.isclose(avg, float(c["penalty_avg"]), rtol=1e-6)
    cl.set_grads(m, pen)
    assert [n for n, p in m.named_parameters() if p.grad is not None] == [str(x) for x in c["get_grads_names"]]
    assert list(cl.get_grads(m).keys()) == [str(x) for x in c["get_grads_names"]]
    # MAS penalty value + autograd gradient
    for p in m.parameters():
        p.grad = None
    val = cl.penalty(m, fish, star)
    val.backward()
    assert np.isclose(val.item(), float(c["mas_penalty"]), rtol=1e-6)
    for n, p in m.named_parameters():
        if p.requires_grad:
            assert np.allclose(p.grad.cpu().numpy(), c["masgrad." + n], rtol=1e-5, atol=1e-8)


def test_plain_dicts_are_accepted():
    """The reference's hooks take ordinary dicts of separate tensors; those are packed, not rejected."""
    m = small_model()
    cur = {n: p.data.clone() for n, p in m.named_parameters() if p.requires_grad}
    ck = {k: v + 0.1 for k, v in cur.items()}
    fi = {k: torch.rand_like(v) for k, v in cur.items()}
    pen, avg = cl.get_penalty_grads({"cl_config": {"e_lambda": 3.0}}, fi, cur, ck)
    o_pen, o_avg = cl_oracle.get_penalty_grads(3.0, {k: v.cpu() for k, v in fi.items()},
                                               {k: v.cpu() for k, v in cur.items()}, {k: v.cpu() for k, v in ck.items()})
    for k in cur:
        assert np.array_equal(pen[k].cpu().numpy(), o_pen[k].numpy())
    assert np.isclose(avg, o_avg, rtol=1e-6)


def test_full_ewc_task_cycle_vs_oracle():
    """penalty-grad injection -> backward on top -> Fisher accumulate over batches -> finalise/merge -> snapshot,
    twice (first task: main = F; second: main = gamma*main + F), against cl_oracle on CPU."""
    m = small_model()
    mc = small_model().cpu()
    mc.load_state_dict({k: v.cpu() for k, v in m.state_dict().items()})
    main_f, o_main = None, None
    gamma = 0.7
    for task in range(2):
        fish = cl.get_zero_params(m, DEV)
        o_fish = {n: torch.zeros_like(p) for n, p in mc.named_parameters() if p.requires_grad}
        total = 0
        for step in range(3):
            x = torch.randn(4 + step, 7)
            for mod, dev in ((m, DEV), (mc, "cpu")):
                for p in mod.parameters():
                    p.grad = None
                loss = mod(x.to(dev)).pow(2).sum(1).mean()
                loss.backward()
                if dev == DEV:
                    cl.fisher_accumulate(fish, cl.get_grads(m), loss)
                else:
                    cl_oracle.fisher_accumulate(o_fish, {n: p.grad for n, p in mc.named_parameters() if p.grad is not None}, loss)
            total += x.shape[0]
        main_f = cl.fisher_finalise(fish, main_f, total, gamma)
        o_main = cl_oracle.fisher_finalise(o_fish, o_main, total, gamma)
        for n in o_main:
            assert rel_err(main_f[n].cpu().numpy(), o_main[n].numpy()) <= 1e-6, (task, n)
        star = cl.get_params_clone(m)
        for n, p in m.named_parameters():
            if p.requires_grad:
                assert torch.equal(star[n], p.data) and star[n].data_ptr() != p.data.data_ptr()
        with torch.no_grad():   # "train": move the parameters
            for (n, p), (_, pc) in zip(m.named_parameters(), mc.named_parameters()):
                if p.requires_grad:
                    d = 0.05 * torch.randn_like(pc)
                    p.add_(d.to(DEV)); pc.add_(d)


def test_mas_cycle_vs_oracle():
    m = small_model()
    imp = cl.get_zero_params(m, DEV)
    o_imp = {n: torch.zeros_like(p).cpu() for n, p in m.named_parameters() if p.requires_grad}
    for step in range(4):
        for p in m.parameters():
            p.grad = None
        m(torch.randn(5, 7, device=DEV)).pow(2).sum(-1).mean().backward()
        cl.mas_accumulate(imp, m)
        cl_oracle.mas_accumulate(o_imp, {n: p.grad.cpu() for n, p in m.named_parameters() if p.grad is not None})
    imp = cl.mas_finalise(imp, 4)
    o_imp = cl_oracle.mas_finalise(o_imp, 4)
    for n in o_imp:
        assert np.allclose(imp[n].cpu().numpy(), o_imp[n].numpy(), rtol=1e-6, atol=0)
    star = cl.get_params_clone(m)
    with torch.no_grad():
        for p in m.parameters():
            if p.requires_grad:
                p.add_(0.1)
    # fast path: value + gradient added into the flat grad buffer in one sweep
    fp = cl.flat_params(m)
    fp.bind_grads(zero=True)
    val = cl.penalty_into_grads(m, imp, star, mas_lambda=2.0)
    th = {n: p.detach().cpu().clone().requires_grad_(True) for n, p in m.named_parameters() if p.requires_grad}
    o_val = cl_oracle.mas_penalty(th, o_imp, {n: v.cpu() for n, v in star.items()})
    (2.0 * o_val).backward()
    assert np.isclose(val.item(), o_val.item(), rtol=1e-6)
    for n, p in m.named_parameters():
        if p.requires_grad:
            assert np.allclose(p.grad.cpu().numpy(), th[n].grad.numpy(), rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("n", [1, 3, 4, 8191, 8192, 8193, 1_000_003])
def test_sweep_sizes_and_tails(n):
    """Ragged sizes (not multiples of 4 / of the chunk) against plain torch elementwise ops, bit-exact."""
    from indic_cl_asr_b200 import _lib
    from indic_cl_asr_b200.cl.flat import FlatDict, Layout

    lay = Layout([("p", torch.Size([n]))])
    g = torch.Generator(device=DEV).manual_seed(n)
    mk = lambda: torch.randn(lay.total, device=DEV, generator=g)
    theta, star, F, grad = FlatDict(lay, mk()), FlatDict(lay, mk()), FlatDict(lay, mk().abs()), mk()
    pen, avg = cl.get_penalty_grads_async(4.0, F, theta, star)
    want = 4.0 * 2 * F["p"] * (theta["p"] - star["p"])
    assert torch.equal(pen["p"], want)
    assert np.isclose(avg.item(), want.abs().mean().item(), rtol=1e-5)
    L = _lib.lib()
    s = _lib.stream_ptr()
    acc = F.flat.clone()
    w = torch.tensor([0.37], device=DEV)
    _lib.check(L.clasr_cl_fisher_accum(acc.data_ptr(), grad.data_ptr(), n, w.data_ptr(), s))
    assert torch.equal(acc[:n], (F.flat[:n] + w * grad[:n] ** 2))
    acc = F.flat.clone()
    _lib.check(L.clasr_cl_mas_accum(acc.data_ptr(), grad.data_ptr(), n, s))
    assert torch.equal(acc[:n], F.flat[:n] + grad[:n].abs())
    dst, src = theta.flat.clone(), F.flat.clone()
    _lib.check(L.clasr_cl_scale_merge(dst.data_ptr(), src.data_ptr(), n, 13.0, 0.5, 0, s))
    assert torch.equal(src[:n], F.flat[:n] / 13.0) and torch.equal(dst[:n], theta.flat[:n] * 0.5 + F.flat[:n] / 13.0)


def test_cl_state_roundtrip(tmp_path):
    """save_cl_state / load_cl_state: theta*, importance and scalars survive a restart bit-for-bit; a model with a
    different trainable-parameter layout is rejected."""
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(13, 7), torch.nn.Linear(7, 5)).to(DEV)
    star = cl.get_params_clone(m)
    imp = cl.get_zero_params(m, DEV)
    imp.flat.uniform_(0, 1)
    p = tmp_path / "cl_state.pt"
    cl.save_cl_state(p, checkpoint=star, importance=imp, lang_idx=3, total_ds=1234)
    m2 = torch.nn.Sequential(torch.nn.Linear(13, 7), torch.nn.Linear(7, 5)).to(DEV)
    star2, imp2, extra = cl.load_cl_state(p, m2)
    assert extra == {"lang_idx": 3, "total_ds": 1234}
    assert torch.equal(star2.flat, star.flat) and torch.equal(imp2.flat, imp.flat)
    assert list(star2.keys()) == list(star.keys()) and star2.is_intact()
    bad = torch.nn.Sequential(torch.nn.Linear(13, 8), torch.nn.Linear(8, 5)).to(DEV)
    with pytest.raises(ValueError):
        cl.load_cl_state(p, bad)
