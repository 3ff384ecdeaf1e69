"""Freeze golden fixtures by running the UNMODIFIED reference in the authoring container.

    python -m oracle.gen_golden          # writes tests/golden/*.npz

TEST INFRASTRUCTURE.  Needs /root/reference (absent on the GPU box, where the committed
fixtures are used instead).  Everything here goes through oracle/ref_import.py.
"""
from __future__ import annotations

import os
import types

import numpy as np
import torch

from . import ref_import as R

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _run_ref_rnnt(acts, labels, act_lens, label_lens, blank, fastemit_lambda=0.0, clamp=-1.0):
    RNNTLossNumba, _ = R.load_rnnt_loss()
    a = torch.tensor(acts, dtype=torch.float32, requires_grad=True)
    fn = RNNTLossNumba(blank=blank, reduction="none", fastemit_lambda=fastemit_lambda, clamp=clamp)
    costs = fn(a, torch.LongTensor(labels), torch.LongTensor(act_lens), torch.LongTensor(label_lens))
    costs.sum().backward()
    return costs.detach().numpy(), a.grad.numpy()


def gen_kat():
    np.savez(os.path.join(OUT, "ref_kat.npz"), **R.kat_literals())


def gen_rnnt():
    out = {}
    rng = np.random.RandomState(0)
    cases = {}
    # reference test_case_small_random (:138-163) and test_case_large_random (:323-352) inputs
    cases["small_random"] = dict(acts=rng.randn(1, 4, 3, 3).astype(np.float32), labels=[[1, 2]], blank=0)
    rng = np.random.RandomState(0)
    cases["large_random"] = dict(
        acts=rng.randn(4, 8, 11, 5).astype(np.float32),
        labels=[[1, 2, 4, 3, 2, 2, 1, 1, 1, 1], [3, 2, 2, 3, 4, 1, 1, 1, 1, 1], [4, 4, 1, 2, 1, 3, 4, 3, 1, 2],
                [1, 1, 2, 1, 2, 3, 3, 1, 1, 1]],
        blank=0,
    )
    rng = np.random.RandomState(7)
    # ragged, blank last (the configuration the model uses: blank = num_classes)
    cases["ragged_blank_last"] = dict(
        acts=(2.0 * rng.randn(5, 9, 6, 8)).astype(np.float32),
        labels=rng.randint(0, 7, size=(5, 5)).tolist(), blank=7,
        act_lens=[9, 7, 4, 1, 9], label_lens=[5, 3, 0, 2, 1],
    )
    rng = np.random.RandomState(11)
    cases["fastemit"] = dict(acts=rng.randn(2, 6, 4, 5).astype(np.float32), labels=[[1, 2, 3], [4, 4, 1]], blank=0,
                             fastemit_lambda=0.01)
    cases["fastemit_clamp"] = dict(acts=rng.randn(2, 6, 4, 5).astype(np.float32), labels=[[1, 2, 3], [2, 2, 1]],
                                   blank=0, fastemit_lambda=0.25, clamp=0.05)
    rng = np.random.RandomState(13)
    cases["wide_vocab"] = dict(acts=(3.0 * rng.randn(2, 5, 4, 300)).astype(np.float32),
                               labels=rng.randint(0, 299, size=(2, 3)).tolist(), blank=299,
                               act_lens=[5, 3], label_lens=[3, 2])
    for name, c in cases.items():
        acts = c["acts"]
        labels = np.asarray(c["labels"], dtype=np.int64)
        act_lens = np.asarray(c.get("act_lens", [acts.shape[1]] * acts.shape[0]), dtype=np.int64)
        label_lens = np.asarray(c.get("label_lens", [labels.shape[1]] * acts.shape[0]), dtype=np.int64)
        fe, cl = c.get("fastemit_lambda", 0.0), c.get("clamp", -1.0)
        costs, grads = _run_ref_rnnt(acts, labels, act_lens, label_lens, c["blank"], fe, cl)
        for k, v in dict(acts=acts, labels=labels, act_lens=act_lens, label_lens=label_lens,
                         blank=np.int64(c["blank"]), fastemit_lambda=np.float64(fe), clamp=np.float64(cl),
                         costs=costs, grads=grads).items():
            out[f"{name}__{k}"] = v
    np.savez(os.path.join(OUT, "ref_rnnt.npz"), **out)


def _joint_case(activation, seed, B, T, U, De, Dp, H, V, fbs, reduction, ragged=True, loss_kwargs=None,
                language_keys=None, language_ids=None):
    """One run of the reference's RNNTJoint (fused branch) + RNNTLoss facade.  ``language_keys``: the multilingual
    ("multisoftmax") joint, V classes PER LANGUAGE: num_classes = V * n_lang, one Linear(H, V+1) per language
    (modules/rnnt.py:1694-1703), loss blank = V (hybrid_rnnt_ctc_bpe_models.py:112-116)."""
    J = R.load_joint_class()
    L = R.load_rnnt_loss_facade()
    torch.manual_seed(seed)
    multi = language_keys is not None
    j = J(jointnet=dict(encoder_hidden=De, pred_hidden=Dp, joint_hidden=H, activation=activation, dropout=0.0),
          num_classes=V * len(language_keys) if multi else V, fuse_loss_wer=True, fused_batch_size=fbs,
          multilingual=multi, language_keys=language_keys)
    loss = L(num_classes=V, reduction=reduction, loss_kwargs=loss_kwargs)
    j.set_loss(loss)
    j.set_wer(object())
    enc = torch.randn(B, De, T, requires_grad=True)
    dec = torch.randn(B, Dp, U + 1, requires_grad=True)
    g = torch.Generator().manual_seed(seed + 1)
    if ragged:
        el = torch.randint(max(1, T // 2), T + 1, (B,), generator=g)
        tl = torch.randint(0, U + 1, (B,), generator=g)
        el[0], tl[0] = T, U
    else:
        el, tl = torch.full((B,), T), torch.full((B,), U)
    tr = torch.randint(0, V, (B, U), generator=g)
    # non-fused logits (fuse flag off -> plain joint), reference forward :1394-1401
    j._fuse_loss_wer = False
    logits = j(encoder_outputs=enc, decoder_outputs=dec, language_ids=language_ids).detach().clone()
    j._fuse_loss_wer = True
    j.store_sub_logits = True
    j.detach_sub_enc = True
    val, _, _, _ = j(encoder_outputs=enc, decoder_outputs=dec, encoder_lengths=el, transcripts=tr,
                     transcript_lengths=tl, compute_wer=False, language_ids=language_ids)
    sub_shapes = np.asarray([list(s.shape) for s in j.store_list], dtype=np.int64)
    val.backward()
    out = dict(
        enc=enc.detach().numpy(), dec=dec.detach().numpy(), enc_lens=el.numpy(), transcripts=tr.numpy(),
        transcript_lens=tl.numpy(), logits=logits.numpy(), loss=val.detach().numpy(),
        d_enc=enc.grad.numpy(), d_dec=dec.grad.numpy(), sub_shapes=sub_shapes,
        cfg=np.asarray([B, T, U, De, Dp, H, V, fbs], dtype=np.int64),
    )
    names = {"enc.weight": j.enc.weight, "enc.bias": j.enc.bias, "pred.weight": j.pred.weight, "pred.bias": j.pred.bias}
    if multi:
        for lang in language_keys:
            names[f"out.{lang}.weight"] = j.joint_net[-1][lang].weight
            names[f"out.{lang}.bias"] = j.joint_net[-1][lang].bias
        out["language_keys"] = np.asarray(language_keys)
        out["language_ids"] = np.asarray(language_ids)
    else:
        names.update({"out.weight": j.joint_net[-1].weight, "out.bias": j.joint_net[-1].bias})
    for k, p in names.items():
        out["p." + k] = p.detach().numpy()
        # a head no utterance of the batch used keeps grad None in the reference (utils.py:308-310 skips it)
        out["g." + k] = p.grad.numpy() if p.grad is not None else np.zeros_like(p.detach().numpy())
    fe = (loss_kwargs or {}).get("fastemit_lambda", 0.0)
    cl = (loss_kwargs or {}).get("clamp", -1.0)
    out["fastemit_lambda"] = np.float64(fe)
    out["clamp"] = np.float64(cl)
    return out


def gen_joint():
    out = {}
    for name, kw in {
        "tanh": dict(activation="tanh", seed=3, B=5, T=7, U=4, De=12, Dp=10, H=16, V=9, fbs=2, reduction="mean_batch"),
        "relu": dict(activation="relu", seed=4, B=4, T=6, U=5, De=8, Dp=8, H=24, V=17, fbs=4, reduction="mean_batch"),
        "sigmoid": dict(activation="sigmoid", seed=5, B=3, T=5, U=3, De=6, Dp=7, H=8, V=5, fbs=1, reduction="mean_volume"),
        "tanh_wide": dict(activation="tanh", seed=6, B=2, T=9, U=6, De=16, Dp=16, H=64, V=40, fbs=4, reduction="sum",
                          ragged=False),
    }.items():
        for k, v in _joint_case(**kw).items():
            out[f"{name}__{k}"] = v
        out[f"{name}__activation"] = np.asarray(kw["activation"])
        out[f"{name}__reduction"] = np.asarray(kw["reduction"])
    np.savez(os.path.join(OUT, "ref_joint.npz"), **out)


def gen_joint_fused():
    """Reference runs whose joint_hidden is a multiple of 64, i.e. shapes the fused tcgen05 kernel accepts: the shipped
    checkpoint's ReLU, the multilingual per-language head with language_ids (one language per batch = the drivers'
    case, and a mixed batch = the per-sample loop :1635-1639), FastEmit, and every NeMo-side reduction."""
    out = {}
    langs = ["hi", "bn", "ta"]
    for name, kw in {
        "relu_h64": dict(activation="relu", seed=51, B=5, T=11, U=6, De=20, Dp=12, H=64, V=29, fbs=2,
                         reduction="mean_batch"),
        "relu_h128_fastemit": dict(activation="relu", seed=52, B=4, T=9, U=5, De=16, Dp=24, H=128, V=33, fbs=4,
                                   reduction="mean_batch", loss_kwargs=dict(fastemit_lambda=0.01)),
        "tanh_h128_mean_volume": dict(activation="tanh", seed=53, B=6, T=10, U=7, De=12, Dp=12, H=128, V=21, fbs=4,
                                      reduction="mean_volume"),
        "sigmoid_h64_mean": dict(activation="sigmoid", seed=54, B=3, T=8, U=4, De=8, Dp=8, H=64, V=12, fbs=2,
                                 reduction="mean", ragged=False),
        # no clamp case here: on CPU the reference joint log-softmaxes its output (:1651-1655) and RNNTLossNumba then
        # clamps the gradient w.r.t. those LOG-PROBS behind a second log_softmax (rnnt_pytorch.py:428-433), which is not
        # what its CUDA path (clamp on the softmax-fused logits gradient, gpu_rnnt_kernel.py:396) computes; clamp is
        # pinned at the loss level instead (ref_kat.npz rnnt_clamp_*, ref_rnnt.npz fastemit_clamp)
        "tanh_h64_fastemit_sum": dict(activation="tanh", seed=55, B=3, T=8, U=5, De=8, Dp=8, H=64, V=15, fbs=3,
                                      reduction="sum", loss_kwargs=dict(fastemit_lambda=0.1)),
        "multilingual_relu": dict(activation="relu", seed=56, B=5, T=9, U=5, De=16, Dp=16, H=64, V=12, fbs=2,
                                  reduction="mean_batch", language_keys=langs, language_ids=["bn"] * 5),
        "multilingual_mixed": dict(activation="tanh", seed=57, B=4, T=7, U=4, De=12, Dp=12, H=64, V=10, fbs=4,
                                   reduction="mean_batch", language_keys=langs,
                                   language_ids=["hi", "ta", "ta", "bn"]),
    }.items():
        for k, v in _joint_case(**kw).items():
            out[f"{name}__{k}"] = v
        out[f"{name}__activation"] = np.asarray(kw["activation"])
        out[f"{name}__reduction"] = np.asarray(kw["reduction"])
    np.savez(os.path.join(OUT, "ref_joint_fused.npz"), **out)


def gen_ctc():
    C = R.load_ctc_loss_class()
    out = {}
    for name, (seed, B, T, U, V, red, zi) in {
        "mean_batch": (21, 4, 12, 4, 6, "mean_batch", True),
        "mean_volume": (22, 3, 10, 3, 5, "mean_volume", True),
        "infeasible": (23, 3, 5, 4, 4, "mean_batch", True),  # T < U+repeats for some samples -> zero_infinity
        "repeats": (24, 2, 14, 6, 3, "mean_batch", True),
    }.items():
        g = torch.Generator().manual_seed(seed)
        logits = (2.0 * torch.randn(B, T, V + 1, generator=g)).requires_grad_(True)
        targets = torch.randint(0, V, (B, U), generator=g)
        il = torch.randint(max(1, T // 2), T + 1, (B,), generator=g)
        tl = torch.randint(0, U + 1, (B,), generator=g)
        il[0], tl[0] = T, U
        if name == "infeasible":
            il[1], tl[1] = 2, U
        lp = logits.log_softmax(-1)
        lp.retain_grad()
        loss = C(num_classes=V, zero_infinity=zi, reduction=red)(
            log_probs=lp, targets=targets, input_lengths=il, target_lengths=tl)
        per = torch.nn.functional.ctc_loss(lp.detach().transpose(0, 1), targets, il, tl, blank=V, reduction="none",
                                           zero_infinity=zi)
        loss.backward()
        for k, v in dict(logits=logits.detach().numpy(), targets=targets.numpy(), input_lens=il.numpy(),
                         target_lens=tl.numpy(), loss=loss.detach().numpy(), per_sample=per.numpy(),
                         d_logits=logits.grad.numpy(), d_log_probs=lp.grad.numpy(),
                         num_classes=np.int64(V), reduction=np.asarray(red)).items():
            out[f"{name}__{k}"] = v
    np.savez(os.path.join(OUT, "ref_ctc.npz"), **out)


def gen_cl():
    H = R.load_cl_hooks()
    torch.manual_seed(31)
    model = torch.nn.Sequential()
    model.add_module("a", torch.nn.Linear(7, 5))
    model.add_module("frozen", torch.nn.Linear(5, 5))
    model.add_module("b", torch.nn.Linear(5, 3, bias=False))
    for p in model.frozen.parameters():
        p.requires_grad = False
    cur = H["get_params"](model)
    ckpt = {k: v + 0.05 * torch.randn_like(v) for k, v in cur.items()}
    fish = {k: torch.rand_like(v) for k, v in cur.items()}
    cfg = types.SimpleNamespace(cl_config=types.SimpleNamespace(e_lambda=10))
    pen, avg = H["get_penalty_grads"](cfg, fish, cur, ckpt)
    mas = H["penalty"](model, fish, ckpt)
    mas.backward()
    out = {"names": np.asarray(list(cur.keys())), "penalty_avg": np.float64(avg), "e_lambda": np.float64(10),
           "mas_penalty": mas.detach().numpy()}
    for k in cur:
        out["theta." + k] = cur[k].numpy().copy()
        out["star." + k] = ckpt[k].numpy()
        out["F." + k] = fish[k].numpy()
        out["pen." + k] = pen[k].numpy()
    for n, p in model.named_parameters():
        if p.requires_grad:
            out["masgrad." + n] = p.grad.numpy().copy()
    # set_grads / get_grads / get_zero_params semantics
    H["set_grads"](model, pen)
    got = H["get_grads"](model)
    out["get_grads_names"] = np.asarray(list(got.keys()))
    z = H["get_zero_params"](model, "cpu")
    out["zero_names"] = np.asarray(list(z.keys()))
    np.savez(os.path.join(OUT, "ref_cl.npz"), **out)


def gen_conv_asr():
    D = R.load_conv_asr_decoder_forward()
    torch.manual_seed(41)
    n_lang, per, De, B, T = 3, 4, 6, 2, 5
    ncls = n_lang * per + 1
    m = D()
    m.decoder_layers = torch.nn.Sequential(torch.nn.Conv1d(De, ncls, kernel_size=1, bias=True))
    m.temperature = 1.0
    m.return_logits_ = True
    masks = {}
    for i, lang in enumerate(["hi", "bn", "ta"]):
        mk = [False] * ncls
        for c in range(i * per, (i + 1) * per):
            mk[c] = True
        mk[-1] = True
        masks[lang] = mk
    m.language_masks = masks
    x = torch.randn(B, De, T)
    lp = m(x, language_ids=["bn", "bn"])
    np.savez(os.path.join(OUT, "ref_conv_asr.npz"), x=x.numpy(), weight=m.decoder_layers[0].weight.detach().numpy(),
             bias=m.decoder_layers[0].bias.detach().numpy(), mask=np.asarray(masks["bn"]),
             log_probs=lp.detach().numpy(), logits=m.decoder_logits.detach().numpy())


def gen_decoder():
    """Reference RNNTDecoder runs (embedding -> zero SOS step -> LSTM): outputs and parameter gradients for a given
    state_dict, so that the work-alike can be checked after load_state_dict (initialisation RNG order is not pinned)."""
    D = R.load_decoder_class()
    out = {}
    for name, (seed, B, U, V, H, L, kw) in {
        "one_layer": (61, 3, 5, 11, 16, 1, {}),
        "two_layers_fgb": (62, 2, 7, 9, 24, 2, dict(forget_gate_bias=0.5, weights_init_scale=0.7,
                                                    hidden_hidden_bias_scale=0.3)),
    }.items():
        torch.manual_seed(seed)
        m = D(prednet=dict(pred_hidden=H, pred_rnn_layers=L, dropout=0.0, **kw), vocab_size=V)
        tg = torch.randint(0, V, (B, U))
        tg[-1, U - 2:] = V          # blank-as-pad positions embed to zero (padding_idx)
        tl = torch.full((B,), U)
        g, tl_out, states = m(targets=tg, target_length=tl)
        w = torch.randn_like(g)
        (g * w).sum().backward()
        out[f"{name}__targets"] = tg.numpy()
        out[f"{name}__g"] = g.detach().numpy()
        out[f"{name}__w"] = w.numpy()
        out[f"{name}__h"] = states[0].detach().numpy()
        out[f"{name}__c"] = states[1].detach().numpy()
        out[f"{name}__cfg"] = np.asarray([B, U, V, H, L], dtype=np.int64)
        for k, p in m.named_parameters():
            out[f"{name}__p.{k}"] = p.detach().numpy()
            out[f"{name}__g.{k}"] = p.grad.numpy()
    np.savez(os.path.join(OUT, "ref_decoder.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    gen_kat()
    gen_rnnt()
    gen_joint()
    gen_joint_fused()
    gen_ctc()
    gen_cl()
    gen_conv_asr()
    gen_decoder()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
