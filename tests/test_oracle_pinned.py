"""The oracle is only trusted after it reproduces (a) the reference's own known-answer vectors and
(b) outputs of the reference itself (frozen by oracle/gen_golden.py in the authoring container)."""
import numpy as np
import pytest
import torch

from oracle import cl_oracle, ctc_oracle, joint_oracle, rnnt_oracle
from conftest import split_cases


def _lens(acts, labels):
    B = acts.shape[0]
    return np.full(B, acts.shape[1], np.int64), np.full(B, labels.shape[1], np.int64)


def test_rnnt_kat_small(golden):
    k = golden("ref_kat.npz")
    acts, labels = k["rnnt_small_acts"], k["rnnt_small_labels"]
    al, ll = _lens(acts, labels)
    costs, grads = rnnt_oracle.rnnt_loss_and_grad(acts, labels, al, ll, blank=0)
    # reference test_rnnt_pytorch.py:95-97,126-127 tolerances
    assert np.allclose(costs.sum(), k["rnnt_small_expected_cost"], atol=1e-6, rtol=1e-6)
    assert np.allclose(grads, k["rnnt_small_expected_grads"], atol=1e-7, rtol=1e-5)


def test_rnnt_kat_big(golden):
    k = golden("ref_kat.npz")
    acts, labels = k["rnnt_big_activations"], k["rnnt_big_labels"]
    al, ll = _lens(acts, labels)
    costs, grads = rnnt_oracle.rnnt_loss_and_grad(acts, labels, al, ll, blank=0)
    assert np.allclose(costs, k["rnnt_big_expected_costs"], atol=1e-7)
    assert np.allclose(grads, k["rnnt_big_expected_grads"], atol=1e-7, rtol=1e-5)


def test_rnnt_kat_clamp(golden):
    k = golden("ref_kat.npz")
    acts, labels = k["rnnt_clamp_acts"], k["rnnt_clamp_labels"]
    al, ll = _lens(acts, labels)
    costs, grads = rnnt_oracle.rnnt_loss_and_grad(acts, labels, al, ll, blank=0, clamp=float(k["rnnt_clamp_GRAD_CLAMP"]))
    assert np.allclose(costs.sum(), k["rnnt_clamp_expected_cost"], atol=1e-6)
    assert np.allclose(grads, k["rnnt_clamp_expected_grads"], atol=1e-7, rtol=1e-5)


def test_rnnt_fastemit_cost_identity(golden):
    # reference test_rnnt_pytorch.py:435-438: cost * (1 + lambda)
    k = golden("ref_kat.npz")
    acts, labels = k["rnnt_small_acts"], k["rnnt_small_labels"]
    al, ll = _lens(acts, labels)
    for lam in (1.0, 0.01, 1e-5):
        costs, _ = rnnt_oracle.rnnt_loss_and_grad(acts, labels, al, ll, blank=0, fastemit_lambda=lam)
        assert np.allclose(costs.sum(), 4.495666 * (1 + lam), rtol=1e-6)


@pytest.mark.parametrize("case", ["small_random", "large_random", "ragged_blank_last", "fastemit", "fastemit_clamp",
                                  "wide_vocab"])
def test_rnnt_vs_reference_run(golden, case):
    c = split_cases(golden("ref_rnnt.npz"))[case]
    clamp = float(c["clamp"])
    costs, grads = rnnt_oracle.rnnt_loss_and_grad(
        c["acts"], c["labels"], c["act_lens"], c["label_lens"], int(c["blank"]),
        fastemit_lambda=float(c["fastemit_lambda"]), clamp=clamp if clamp > 0 else 0.0)
    assert np.allclose(costs, c["costs"], rtol=1e-5, atol=1e-5)
    # stated tolerance: 1e-4 relative on gradients (reference runs fp32; oracle is fp64)
    assert np.abs(grads - c["grads"]).max() <= 1e-4 * np.abs(c["grads"]).max()


def test_alpha_beta_diag_matches_loops():
    rng = np.random.RandomState(3)
    T, U1 = 9, 6
    lb, ll = -np.abs(rng.randn(T, U1)), -np.abs(rng.randn(T, U1))
    a1 = rnnt_oracle.alphas_betas(lb, ll, T, U1)
    a2 = rnnt_oracle.alphas_betas_diag(lb, ll, T, U1)
    for x, y in zip(a1, a2):
        assert np.allclose(x, y)
    assert abs(a1[2] - a1[3]) < 1e-9  # llForward == llBackward (cpu_rnnt.py:240-242)


@pytest.mark.parametrize("name", ["small", "blank_last"])
def test_ctc_kat(golden, name):
    k = golden("ref_kat.npz")
    logits = k[f"ctc_{name}_acts"]
    labels = k[f"ctc_{name}_labels"]
    blank = 0 if name == "small" else logits.shape[-1] - 1
    lp = logits - np.log(np.exp(logits).sum(-1, keepdims=True))
    nll, g = ctc_oracle.ctc_loss_and_grad(lp, labels, [logits.shape[1]], [labels.shape[1]], blank)
    assert np.allclose(nll.sum(), k[f"ctc_{name}_expected_cost"], rtol=1e-6)
    # goldens are w.r.t. logits: compose with log_softmax backward
    gl = g - np.exp(lp) * g.sum(-1, keepdims=True)
    assert np.allclose(gl, k[f"ctc_{name}_expected_grads"], atol=1e-6)


@pytest.mark.parametrize("case", ["mean_batch", "mean_volume", "infeasible", "repeats"])
def test_ctc_vs_reference_run(golden, case):
    c = split_cases(golden("ref_ctc.npz"))[case]
    logits = c["logits"].astype(np.float64)
    lp = logits - np.log(np.exp(logits - logits.max(-1, keepdims=True)).sum(-1, keepdims=True)) - logits.max(-1, keepdims=True)
    V = int(c["num_classes"])
    nll, g = ctc_oracle.ctc_loss_and_grad(lp, c["targets"], c["input_lens"], c["target_lens"], V, zero_infinity=True)
    assert np.allclose(nll, c["per_sample"], rtol=1e-5, atol=1e-5)
    red = str(c["reduction"])
    assert np.allclose(ctc_oracle.reduce_losses(nll, c["target_lens"], red), c["loss"], rtol=1e-5)
    scale = 1.0 / len(nll) if red == "mean_batch" else 1.0 / c["target_lens"].sum()
    assert np.allclose(g * scale, c["d_log_probs"], atol=2e-6)
    gl = g - np.exp(lp) * g.sum(-1, keepdims=True)
    assert np.allclose(gl * scale, c["d_logits"], atol=2e-6)


@pytest.mark.parametrize("case", ["tanh", "relu", "sigmoid", "tanh_wide"])
def test_joint_vs_reference_run(golden, case):
    c = split_cases(golden("ref_joint.npz"))[case]
    B, T, U, De, Dp, H, V, fbs = [int(x) for x in c["cfg"]]
    act, red = str(c["activation"]), str(c["reduction"])
    p = {k[2:]: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in c.items() if k.startswith("p.")}
    enc = torch.tensor(c["enc"], dtype=torch.float64, requires_grad=True)
    dec = torch.tensor(c["dec"], dtype=torch.float64, requires_grad=True)
    z = joint_oracle.joint_logits(enc.transpose(1, 2), dec.transpose(1, 2), p, act)
    # the reference joint log-softmaxes CPU tensors (modules/rnnt.py:1651-1655), so the fixture holds log-probs
    assert np.allclose(z.log_softmax(-1).detach().numpy(), c["logits"], atol=2e-6)
    loss = joint_oracle.fused_joint_loss(enc, dec, torch.tensor(c["enc_lens"]), torch.tensor(c["transcripts"]),
                                         torch.tensor(c["transcript_lens"]), p, act, V, fbs, red)
    assert np.allclose(loss.item(), c["loss"], rtol=1e-5)
    loss.backward()
    for k, t in p.items():
        ref = c["g." + k]
        assert np.abs(t.grad.numpy() - ref).max() <= 1e-4 * max(np.abs(ref).max(), 1e-3), k
    assert np.abs(enc.grad.numpy() - c["d_enc"]).max() <= 1e-4 * np.abs(c["d_enc"]).max()
    assert np.abs(dec.grad.numpy() - c["d_dec"]).max() <= 1e-4 * np.abs(c["d_dec"]).max()


FUSED_CASES = ["relu_h64", "relu_h128_fastemit", "tanh_h128_mean_volume", "sigmoid_h64_mean", "tanh_h64_fastemit_sum",
               "multilingual_relu", "multilingual_mixed"]


@pytest.mark.parametrize("case", FUSED_CASES)
def test_joint_fused_shapes_vs_reference_run(golden, case):
    """ref_joint_fused.npz: reference runs at joint_hidden in {64, 128} (the shapes the tcgen05 kernel accepts) —
    ReLU, multilingual head with language_ids (single-language and mixed batches), FastEmit, all reductions."""
    c = split_cases(golden("ref_joint_fused.npz"))[case]
    B, T, U, De, Dp, H, V, fbs = [int(x) for x in c["cfg"]]
    act, red = str(c["activation"]), str(c["reduction"])
    lang = [str(x) for x in c["language_ids"]] if "language_ids" in c else None
    p = {k[2:]: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in c.items() if k.startswith("p.")}
    enc = torch.tensor(c["enc"], dtype=torch.float64, requires_grad=True)
    dec = torch.tensor(c["dec"], dtype=torch.float64, requires_grad=True)
    z = joint_oracle.joint_logits(enc.transpose(1, 2), dec.transpose(1, 2), p, act, lang)
    assert np.allclose(z.log_softmax(-1).detach().numpy(), c["logits"], atol=2e-6)
    loss = joint_oracle.fused_joint_loss(enc, dec, torch.tensor(c["enc_lens"]), torch.tensor(c["transcripts"]),
                                         torch.tensor(c["transcript_lens"]), p, act, V, fbs, red,
                                         fastemit_lambda=float(c["fastemit_lambda"]),
                                         clamp=max(0.0, float(c["clamp"])), language_ids=lang)
    assert np.allclose(loss.item(), c["loss"], rtol=1e-5)
    loss.backward()
    for k, t in p.items():
        ref = c["g." + k]
        got = t.grad.numpy() if t.grad is not None else np.zeros_like(ref)   # a head no utterance used
        assert np.abs(got - ref).max() <= 1e-4 * max(np.abs(ref).max(), 1e-3), k
    assert np.abs(enc.grad.numpy() - c["d_enc"]).max() <= 1e-4 * np.abs(c["d_enc"]).max()
    assert np.abs(dec.grad.numpy() - c["d_dec"]).max() <= 1e-4 * np.abs(c["d_dec"]).max()


def test_cl_vs_reference_run(golden):
    c = golden("ref_cl.npz")
    names = [str(n) for n in c["names"]]
    assert names == ["a.weight", "a.bias", "b.weight"]  # requires_grad-filtered named_parameters order
    theta = {n: torch.tensor(c["theta." + n]) for n in names}
    star = {n: torch.tensor(c["star." + n]) for n in names}
    F = {n: torch.tensor(c["F." + n]) for n in names}
    pen, avg = cl_oracle.get_penalty_grads(float(c["e_lambda"]), F, theta, star)
    for n in names:
        assert np.allclose(pen[n].numpy(), c["pen." + n], rtol=1e-6, atol=1e-8)
    assert np.isclose(avg, float(c["penalty_avg"]), rtol=1e-6)
    th = {n: t.clone().requires_grad_(True) for n, t in theta.items()}
    m = cl_oracle.mas_penalty(th, F, star)
    assert np.isclose(m.item(), float(c["mas_penalty"]), rtol=1e-6)
    m.backward()
    for n in names:
        assert np.allclose(th[n].grad.numpy(), c["masgrad." + n], rtol=1e-5, atol=1e-8)


def test_conv_asr_vs_reference_run(golden):
    c = golden("ref_conv_asr.npz")
    idx = torch.nonzero(torch.tensor(c["mask"])).flatten()
    lp, z = joint_oracle.ctc_head(torch.tensor(c["x"]), torch.tensor(c["weight"]), torch.tensor(c["bias"]), idx)
    assert np.allclose(z.numpy(), c["logits"], atol=1e-6)
    assert np.allclose(lp.numpy(), c["log_probs"], atol=1e-6)


def test_c_port_lattice_vs_kat_and_reference_runs(golden):
    """oracle/lattice.c (the C/OpenMP port of cpu_rnnt.py that bench.py's cpu_baseline times) against the reference's
    own known-answer vectors and reference runs: costs and gradients w.r.t. the logits (autograd through log_softmax,
    exactly the reference's CPU composition, rnnt_pytorch.py:411-437)."""
    from oracle import c_port

    k = golden("ref_kat.npz")
    for name, cost_key, acts_key in (("small", "rnnt_small_expected_cost", "rnnt_small_acts"),
                                     ("big", "rnnt_big_expected_costs", "rnnt_big_activations")):
        acts = torch.tensor(k[acts_key], dtype=torch.float32, requires_grad=True)
        labels = torch.tensor(k[f"rnnt_{name}_labels"], dtype=torch.long)
        B, T = acts.shape[0], acts.shape[1]
        costs = c_port.rnnt_loss_cpu(acts, labels, torch.full((B,), T), torch.full((B,), labels.shape[1]), 0)
        costs.sum().backward()
        assert np.allclose(costs.detach().numpy().sum(), np.asarray(k[cost_key]).sum(), rtol=1e-6)
        assert np.allclose(acts.grad.numpy(), k[f"rnnt_{name}_expected_grads"], atol=1e-6, rtol=1e-4)
    c = split_cases(golden("ref_rnnt.npz"))["ragged_blank_last"]
    acts = torch.tensor(c["acts"], requires_grad=True)
    costs = c_port.rnnt_loss_cpu(acts, torch.tensor(c["labels"]), torch.tensor(c["act_lens"]), torch.tensor(c["label_lens"]),
                                 int(c["blank"]))
    costs.sum().backward()
    assert np.allclose(costs.detach().numpy(), c["costs"], rtol=1e-5, atol=1e-5)
    assert np.allclose(acts.grad.numpy(), c["grads"], atol=2e-6, rtol=1e-4)
