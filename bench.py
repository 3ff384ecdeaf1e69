#!/usr/bin/env python
"""bench.py — RNNT+CTC+EWC fwd/bwd utterances/s at B=32, T=250, U=100, V=1024 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port), rank 0 only
    python bench.py --config 3|4|5 ...                       # the other BASELINE.json configs (parity-test shapes)

One "step" = one pass of the hot path over one batch of synthetic input:
  fused joint (enc/pred projections -> tcgen05 joint GEMM + online log-softmax) -> alpha/beta wavefront -> RNNT cost,
  CTC head (Conv1d k=1 -> log_softmax kernel) -> CTC forward-backward, mixed loss 0.7/0.3, backward to d enc_out,
  d pred_out and every joint / CTC-head parameter, EWC penalty gradient pre-loaded into the flat gradient buffer
  (cl_baseline_ewc.py:228-240), and - for N > 1 - ONE NCCL sum all-reduce of the flat gradient buffer.
The whole step (both streams and the collective) is recorded into a CUDA graph and replayed (indic_cl_asr_b200/graph.py).

Scaling is STRONG: the global batch of 32 utterances is sharded, 32/N per GPU; `value` is
32 * steps / time.  The weak-scaling figure (32 utterances on EVERY GPU) is reported beside it as
`roofline_more.weak_scaling` when N > 1.
The default backward mode recomputes the logits (they never reach HBM, BASELINE.json north_star (1)); the faster
mode that keeps them between forward and backward is measured beside it as `roofline_more.stash_mode`.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

CTC_WEIGHT = 0.3   # cfg.aux_ctc.ctc_loss_weight (hybrid_rnnt_ctc_models.py:233,902)
E_LAMBDA = 10.0    # config.yaml e_lambda

# BASELINE.json configs (index + 1).  Config 2 is the one the metric is quoted on.
CONFIGS = {
    2: dict(workload="configs[1]: standalone RNNT+CTC(+EWC) loss fwd/bwd, B=32 T=250 U=100 V=1024 H=640",
            B=32, T=250, U=100, V=1024, H=640, D_enc=512, D_pred=640, activation="tanh", dropout=0.0,
            precision="auto", ragged=False, n_lang=0, prednet=False),
    # IndicConformer-medium shapes: d_model 256, pred/joint 640, 16 s audio -> T' = 400 after 4x
    # subsampling, 22 languages x 256 tokens (joint: one Linear(640,257) per language; CTC head: 5633 classes with a
    # per-language mask), ReLU joint with dropout 0.2.  Encoder output is synthetic (the Conformer is upstream of the
    # path); the prediction network (embedding + LSTM, cuDNN) runs inside the step.
    3: dict(workload="configs[2]: IndicConformer-medium hybrid RNNT-CTC EWC fine-tune step (loss half + prediction net on "
                     "synthetic encoder output), B=32 T'=400 U<=100 V=256/language x 22, ReLU, dropout 0.2, ragged",
            B=32, T=400, U=100, V=256, H=640, D_enc=256, D_pred=640, activation="relu", dropout=0.2,
            precision="auto", ragged=True, n_lang=22, prednet=True),
    5: dict(workload="configs[4]: Conformer-large hybrid RNNT-CTC loss fwd/bwd, V=4096, B=64 T=500 U=200, bf16 joint GEMM",
            B=64, T=500, U=200, V=4096, H=640, D_enc=512, D_pred=640, activation="tanh", dropout=0.0,
            precision="bf16", ragged=False, n_lang=0, prednet=False),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE.json configs[config-1]")
    ap.add_argument("--mode", default="tcgen05", choices=["tcgen05", "materialised"])
    ap.add_argument("--precision", default=None, choices=["auto", "bf16x3", "fp16x3", "fp16m8", "bf16"],
                    help="default: the config's (auto = the module default, fp16m8)")
    ap.add_argument("--backward", default="recompute", choices=["recompute", "stash"],
                    help="recompute (module default): the logits never reach HBM; stash: keep them for the backward pass")
    ap.add_argument("--activation", default=None, choices=["tanh", "relu", "sigmoid"])
    ap.add_argument("--ragged", action="store_true", help="enc_len ~ U{T/2..T}, tgt_len ~ U{U/2..U} (SURVEY.md §8d)")
    ap.add_argument("--dropout", type=float, default=None, help="joint dropout (the shipped checkpoint trains with 0.2)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--batch", type=int, default=None, help="override the config's global batch (experiments: --batch 4 on one "
                                                            "GPU is the per-GPU work of the 8-GPU strong-scaling point)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    ap.add_argument("--cpu-sample", type=int, default=4, help="utterances in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the stand-alone roofline probes and the stash-mode run")
    ap.add_argument("--no-overlap-ctc", action="store_true", help="keep the CTC branch on the main stream")
    ap.add_argument("--ewc-params", type=int, default=120_000_000, help="size of the stand-alone regulariser sweep probes")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def resolve_cfg(args):
    c = dict(CONFIGS[args.config if args.config != 4 else 2])
    if args.precision is not None:
        c["precision"] = args.precision
    if args.activation is not None:
        c["activation"] = args.activation
    if args.dropout is not None:
        c["dropout"] = args.dropout
    if args.ragged:
        c["ragged"] = True
    if getattr(args, "batch", None):
        c["B"] = args.batch
        c["workload"] += f" [batch overridden: B={args.batch}]"
    return c


def synth(c, B, seed):
    """Seeded synthetic inputs in NeMo layouts: enc [B,D_enc,T], dec [B,D_pred,U+1], transcripts, lengths."""
    g = torch.Generator().manual_seed(seed)
    enc = torch.randn(B, c["D_enc"], c["T"], generator=g)
    dec = torch.randn(B, c["D_pred"], c["U"] + 1, generator=g)
    tr = torch.randint(0, c["V"], (B, c["U"]), generator=g)
    if c["ragged"]:
        el = torch.randint(c["T"] // 2, c["T"] + 1, (B,), generator=g)
        tl = torch.randint(c["U"] // 2, c["U"] + 1, (B,), generator=g)
        el[0], tl[0] = c["T"], c["U"]
    else:
        el, tl = torch.full((B,), c["T"]), torch.full((B,), c["U"])
    return enc, dec, tr, el, tl


# ----------------------------------------------------------------------------------------------- CPU baseline
def cpu_reference_step(c, sample, modules_state=None):
    """The reference's CPU path on a bounded sample: restated joint (torch CPU, all threads) + transducer loss
    (oracle/lattice.c, a C port of cpu_rnnt.py) + torch.nn.CTCLoss + dict-of-tensors EWC (oracle/cl_oracle.py)."""
    from oracle import c_port, cl_oracle, joint_oracle

    enc, dec, tr, el, tl = sample
    B = enc.shape[0]
    torch.manual_seed(0)
    if modules_state is None:
        lin = lambda i, o: (torch.nn.Linear(i, o).weight.detach().requires_grad_(True),
                            torch.nn.Linear(i, o).bias.detach().requires_grad_(True))
        ew, eb = lin(c["D_enc"], c["H"])
        pw, pb = lin(c["D_pred"], c["H"])
        ow, ob = lin(c["H"], c["V"] + 1)
        cw = (0.05 * torch.randn(c["V"] + 1, c["D_enc"], 1)).requires_grad_(True)
        cb = torch.zeros(c["V"] + 1, requires_grad=True)
        params = {"enc.weight": ew, "enc.bias": eb, "pred.weight": pw, "pred.bias": pb, "out.weight": ow,
                  "out.bias": ob, "ctc.weight": cw, "ctc.bias": cb}
        star = {k: (v.detach() + 0.01 * torch.randn_like(v)) for k, v in params.items()}
        fish = {k: torch.rand_like(v) for k, v in params.items()}
        modules_state = (params, star, fish)
    params, star, fish = modules_state
    for p in params.values():
        p.grad = None
    enc = enc.clone().requires_grad_(True)
    dec = dec.clone().requires_grad_(True)
    t0 = time.perf_counter()
    # fused sub-batch loop, fused_batch_size = 4 like the shipped checkpoint (modules/rnnt.py:1425)
    losses = []
    for b0 in range(0, B, 4):
        sl = slice(b0, min(B, b0 + 4))
        z = joint_oracle.joint_logits(enc[sl].transpose(1, 2), dec[sl].transpose(1, 2),
                                      {k: params[k] for k in ("enc.weight", "enc.bias", "pred.weight", "pred.bias",
                                                              "out.weight", "out.bias")}, c["activation"])
        losses.append(c_port.rnnt_loss_cpu(z, tr[sl], el[sl], tl[sl], c["V"]))
    loss_rnnt = torch.cat(losses).mean()
    lp, _ = joint_oracle.ctc_head(enc, params["ctc.weight"], params["ctc.bias"])
    loss_ctc = torch.nn.functional.ctc_loss(lp.transpose(0, 1), tr, el, tl, blank=c["V"], reduction="none",
                                            zero_infinity=True).mean()
    loss = (1 - CTC_WEIGHT) * loss_rnnt + CTC_WEIGHT * loss_ctc
    pen, avg = cl_oracle.get_penalty_grads(E_LAMBDA, fish, {k: v.data for k, v in params.items()}, star)
    for k, p in params.items():
        p.grad = pen[k]
    loss.backward()
    dt = time.perf_counter() - t0
    return dt, float(loss.detach()), modules_state


def run_cpu_baseline(c, n_utts, steps=1, warmup=0):
    torch.set_num_threads(os.cpu_count() or 1)
    cc = dict(c, ragged=False)
    sample = synth(cc, n_utts, 1234)
    state = None
    for _ in range(warmup):
        _, _, state = cpu_reference_step(cc, sample, state)
    times = []
    for _ in range(steps):
        dt, loss, state = cpu_reference_step(cc, sample, state)
        times.append(dt)
    total = sum(times)
    return dict(value=n_utts * steps / total, unit="utts/s", cores=torch.get_num_threads(), kind="port",
                sample=f"{n_utts} utterances of the same shape (T={c['T']},U={c['U']},V={c['V']},H={c['H']}), "
                       f"{steps} step(s): torch-CPU joint fwd/bwd + C port of cpu_rnnt.py + torch CTCLoss + EWC dict sweep",
                seconds=total, ms_per_step=1e3 * total / steps, loss=loss)


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML (10 ms period) when importable, else
    nvidia-smi (the B200_PROFILING.md clocks line, ~5 samples/s)."""

    # NVML clocks-event (throttle) reason bits; gpu_idle (0x1) is not a throttle and is left out
    _REASONS = (("applications_clocks_setting", 0x2), ("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sync_boost", 0x10),
                ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40), ("hw_power_brake_slowdown", 0x80),
                ("display_clock_setting", 0x100))

    def __init__(self, index):
        self.index = index
        self.samples = []   # (sm_mhz, sm_max_mhz, [reasons])
        self._stop = threading.Event()
        self._t = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(int(os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[index])
                                                        if os.environ.get("CUDA_VISIBLE_DEVICES") else index)
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        sm = float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
        try:
            mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self._h))
        except Exception:
            mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
        self.samples.append((sm, self._max, [name for name, bit in self._REASONS if mask & bit]))

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
            p = [x.strip() for x in out.split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            self.samples.append((float(p[0]), float(p[1]),
                                 [nm for nm, v in zip(names, p[2:6]) if v.lower().startswith("active")]))

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample_nvml() if self._nvml else self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.005 if self._nvml else 0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = sorted({r for s in self.samples for r in s[2]})
        return {"sm_mhz": statistics.median(s[0] for s in self.samples), "sm_max_mhz": max(s[1] for s in self.samples),
                "reasons": reasons, "samples": len(self.samples),
                "source": "nvml" if self._nvml else "nvidia-smi"}


# ----------------------------------------------------------------------------------------------- our arm
class Workload:
    """The modules of one config on one device, the EWC state, and the step closure."""

    def __init__(self, c, args, dev, world, backward_mode):
        from indic_cl_asr_b200 import (CTCLoss, ConvASRDecoder, HybridRNNTCTCLoss, RNNTDecoder, RNNTJoint, RNNTLoss, cl)

        self.c, self.dev, self.world = c, dev, world
        torch.manual_seed(1234)
        nl = c["n_lang"]
        keys = [f"l{i:02d}" for i in range(nl)] if nl else None
        total_v = c["V"] * nl if nl else c["V"]
        joint = RNNTJoint(jointnet=dict(encoder_hidden=c["D_enc"], pred_hidden=c["D_pred"], joint_hidden=c["H"],
                                        activation=c["activation"], dropout=c["dropout"]),
                          num_classes=total_v, fuse_loss_wer=True, fused_batch_size=4, fused_impl=args.mode,
                          precision=c["precision"], backward_mode=backward_mode, multilingual=bool(nl),
                          language_keys=keys).to(dev)
        self.precision = joint.precision   # 'auto' resolved by the module
        joint.set_loss(RNNTLoss(num_classes=c["V"]))
        joint.set_wer(object())
        masks = None
        if nl:   # per-language mask over the (n_lang * V + 1)-way CTC head: that language's block + the blank
            masks = {}
            for i, k in enumerate(keys):
                m = [False] * (total_v + 1)
                m[i * c["V"]:(i + 1) * c["V"]] = [True] * c["V"]
                m[-1] = True
                masks[k] = m
        head = ConvASRDecoder(feat_in=c["D_enc"], num_classes=total_v, language_masks=masks).to(dev)
        ctc = CTCLoss(num_classes=c["V"], zero_infinity=True)
        mods = {"joint": joint, "ctc_decoder": head}
        self.decoder = None
        if c["prednet"]:
            self.decoder = RNNTDecoder(prednet=dict(pred_hidden=c["D_pred"], pred_rnn_layers=1, dropout=0.0),
                                       vocab_size=total_v).to(dev)
            mods["decoder"] = self.decoder
        self.lang_ids_of = (lambda B: [keys[3]] * B) if nl else (lambda B: None)
        self.lang_offset = 3 * c["V"] if nl else 0
        self.model = torch.nn.ModuleDict(mods)
        self.joint, self.head = joint, head
        self.hybrid = HybridRNNTCTCLoss(joint, head, ctc, ctc_loss_weight=CTC_WEIGHT, overlap_ctc=not args.no_overlap_ctc)
        self.cl = cl
        self.fp = cl.flat_params(self.model)
        self.theta = cl.get_params(self.model)
        self.star = cl.get_params_clone(self.model)
        self.star.flat.add_(0.01 * torch.randn_like(self.star.flat))
        self.fish = cl.get_zero_params(self.model, dev)
        self.fish.flat.uniform_(0.0, 1.0)
        # every rank writes the penalty gradient into its buffer before the SUM all-reduce: 1/N of it per rank
        self.ewc_cfg = {"cl_config": {"e_lambda": E_LAMBDA / world}}
        self.n_params = int(self.fp.layout.total)

    def make_step(self, B_local, B_global, variant="ewc"):
        """variant 'ewc': training step of a task > 0 (penalty gradient pre-loaded, cl_baseline_ewc.py:228-240);
        'fisher': step of the importance epoch (no penalty, Fisher += loss * grad^2, :245-255)."""
        from indic_cl_asr_b200.dist import allreduce_flat_, local_loss_scale

        scale = local_loss_scale(B_local, B_global)
        lang = self.lang_ids_of(B_local)
        fp, cl, world = self.fp, self.cl, self.world

        pen_stream = torch.cuda.Stream(device=self.dev)
        pen_overlap = os.environ.get("CLASR_PEN_OVERLAP", "0") != "0"   # A/B switch; measured neutral to slower

        def step(enc, dec, tr, el, tl):
            cur = torch.cuda.current_stream(self.dev)
            if variant == "ewc":
                # the penalty sweep only has to land in the gradient buffer before backward() accumulates on top of it
                # (cl_baseline_ewc.py:228-240): it runs beside the forward pass on its own stream
                fp.bind_grads(zero=False)
                if pen_overlap:
                    pen_stream.wait_stream(cur)
                with torch.cuda.stream(pen_stream if pen_overlap else cur):
                    _, avg = cl.get_penalty_grads_async(self.ewc_cfg, self.fish, self.theta, self.star, out=fp.grad)
            else:
                fp.bind_grads(zero=True)
                avg = None
            enc.grad = None
            dec.grad = None
            if self.decoder is not None:   # prediction network: embedding + SOS + LSTM -> [B, D, U+1]
                dec_out, _, _ = self.decoder(targets=tr + self.lang_offset, target_length=tl)
            else:
                dec_out = dec
            loss, _ = self.hybrid(enc, el, dec_out, tr, tl, language_ids=lang)   # hybrid_rnnt_ctc_models.py:868-902
            if variant == "ewc" and pen_overlap:
                cur.wait_stream(pen_stream)
            (loss * scale).backward()      # local mean_batch x B_local/B: the SUM all-reduce gives the global mean's gradient
            if world > 1:
                allreduce_flat_(fp.grad)
            if variant == "fisher":
                cl.fisher_accumulate(self.fish, fp.grad_dict(), loss)
                avg = loss.detach()
            return loss.detach().float().reshape(1), avg.detach().float().reshape(1)

        return step


def time_steps(run, steps, flush, barrier, dev, world):
    """K steps, each bracketed by CUDA events on the launch stream, L2 flushed in between (outside the event pairs);
    returns (total ms = MAX over ranks, per-step list of this rank)."""
    import torch.distributed as dist

    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    for s_, e_ in evs:
        flush.zero_()                      # evict L2 between timed iterations
        s_.record()
        run()
        e_.record()
    barrier()
    per = [s_.elapsed_time(e_) for s_, e_ in evs]
    t = torch.tensor([sum(per)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), per


def traffic_from_profiles(key):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel from the committed ncu capture
    of this round (profiles/r02_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep); None if no capture
    matches this run's configuration."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(path):
        return None, None
    with open(path) as fh:
        d = json.load(fh)
    ent = d.get(key)
    if not ent:
        return None, None
    return float(ent["dram_bytes_read"]) + float(ent["dram_bytes_write"]), ent.get("source")


def main_b200(args):
    import torch.distributed as dist

    from indic_cl_asr_b200 import _lib
    from indic_cl_asr_b200.dist import shard_bounds
    from indic_cl_asr_b200.graph import GraphedStep
    from indic_cl_asr_b200.hybrid import silence_side_stream_grad_warning

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        # a mismatched collective must fail within minutes, not hold the box for NCCL's default 10-minute watchdog
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"),
                                timeout=datetime.timedelta(seconds=120))
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    # "fp32" config: torch's own small GEMMs/convs stay true fp32.  Config 3 runs the prediction network's LSTM (cuDNN,
    # out of scope) inside the step: it keeps torch's DEFAULT cudnn.allow_tf32 = True, as the reference's drivers do —
    # with it off cuDNN issues 101 SIMT-sgemm launches per direction (5.8 of 10.8 ms, profiles/r02e)
    torch.backends.cudnn.allow_tf32 = args.config == 3
    torch.backends.cuda.matmul.allow_tf32 = False
    silence_side_stream_grad_warning()
    if args.config == 4:
        return main_config4(args, dev, world, rank)
    c = resolve_cfg(args)
    pk = peaks()
    L = _lib.lib()
    steps, warm = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B = c["B"]
    strong = args.scaling == "strong"
    if strong:
        b0, b1 = shard_bounds(B, rank, world)
        B_local, B_global = b1 - b0, B
        if B_local == 0:
            raise SystemExit(f"--gpus {world}: nothing left to shard (B={B})")
    else:
        b0, B_local, B_global = 0, B, B * world
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def host_inputs(n, offset, seed):
        full = synth(c, max(B, n + offset), seed)
        return [x[offset:offset + n].contiguous().pin_memory() for x in full]

    def device_inputs(host):
        e, d, tr, el, tl = [x.to(dev) for x in host]
        return [e.requires_grad_(True), d.requires_grad_(True), tr, el, tl]

    # strong scaling: every rank draws the SAME global batch and keeps its slice; weak: its own batch
    host = host_inputs(B_local, b0, 1234) if strong else host_inputs(B_local, 0, 1234 + rank)

    prof_names = ("joint_fwd", "joint_bwd_dz", "joint_dz_sweep", "gemm_dhid", "gemm_dw", "joint_dfg", "rnnt_lattice",
                  "rnnt_lse", "rnnt_grad", "ctc_lattice", "ctc_grad", "cl_penalty_grad")

    def measure(backward_mode, B_loc, B_glob, host_in, variant="ewc", want_e2e=False, want_profile=False):
        """Build the modules, capture the step, time K replays.  Returns a dict of results."""
        wl = Workload(c, args, dev, world, backward_mode)
        step = wl.make_step(B_loc, B_glob, variant)
        ins = device_inputs(host_in)
        n0 = _lib.launch_count()
        for _ in range(warm):
            step(*ins)
        torch.cuda.synchronize()
        launches_per_step = (_lib.launch_count() - n0) // warm
        kern, n_prof = {}, min(steps, 5)
        if want_profile:
            # per-kernel durations (library-side CUDA events on the launch stream), EAGER launches, before the graph
            # takes its memory; the CTC branch stays on the main stream so that no duration includes shared time
            wl.hybrid.overlap_ctc = False
            L.clasr_set_profiling(1)
            L.clasr_profile_reset()
            for _ in range(n_prof):
                flush.zero_()
                step(*ins)
            torch.cuda.synchronize()
            kern = {k: _lib.profile_ms(k) for k in prof_names}
            kern = {k: v for k, v in kern.items() if v >= 0}
            if "rnnt_lse" in kern:   # materialised mode: several launches per step (the reference's sub-batch loop)
                kern["rnnt_lse+grad_per_step"] = sum(
                    m * n for m, n in (_lib.profile_ms_count(k) for k in ("rnnt_lse", "rnnt_grad")) if n > 0) / n_prof
            L.clasr_set_profiling(0)
            wl.hybrid.overlap_ctc = not args.no_overlap_ctc
        gs, graph_error = None, None
        if not args.no_graph:
            torch.cuda.empty_cache()
            try:
                gs = GraphedStep(step, ins, warmup=1)
            except Exception as ex:   # report, then measure the eager step rather than nothing
                graph_error = f"{type(ex).__name__}: {ex}"[:300]
                print(f"[bench] CUDA graph capture failed, falling back to eager launches: {graph_error}", file=sys.stderr)
                torch.cuda.synchronize()
        run = gs.replay if gs is not None else (lambda: step(*ins))
        for _ in range(2):
            run()
        with ClockSampler(local_rank) as clk:
            total_ms, per = time_steps(run, steps, flush, barrier, dev, world)
            # an NVML query can take tens of ms: keep the identical load running (untimed) so that the sampler sees
            # enough of it for a median.  A FIXED count: the step contains a collective, every rank must replay it
            # the same number of times.
            for _ in range(40):
                run()
            torch.cuda.synchronize()
        res = dict(wl=wl, step=step, ins=ins, gs=gs, total_ms=total_ms, per=per, ms=total_ms / steps,
                   value=B_glob * steps / (total_ms / 1e3), launches_per_step=int(launches_per_step),
                   clocks=clk.summary(), precision=wl.precision, n_params=wl.n_params, graph_error=graph_error, kern=kern)
        if want_e2e:
            res["e2e"] = measure_e2e(wl, step, ins, gs, host_in, B_glob)
        return res

    def measure_e2e(wl, step, ins, gs, host_in, B_glob):
        """The same metric through the public API with HOST inputs: every step's inputs cross from pinned host memory
        (one cudaMemcpyAsync per tensor, on a copy stream, prefetched one step ahead into the other of two input
        sets) and the step's loss and penalty_avg are read back; ONE timed region around all K steps."""
        sets = [ins]
        graphs = [gs]
        if gs is not None:   # second input set + its own graph (shared memory pool: the two never run concurrently)
            ins_b = device_inputs(host_in)
            sets.append(ins_b)
            graphs.append(GraphedStep(step, ins_b, warmup=0, pool=gs.pool()))
        main = torch.cuda.current_stream(dev)
        copy = torch.cuda.Stream(device=dev)
        h2d = sum(x.numel() * x.element_size() for x in host_in)
        host_out = torch.empty(steps + warm, 2, dtype=torch.float32).pin_memory()
        n = len(sets)
        h2d_done = [torch.cuda.Event() for _ in range(n)]
        used = [torch.cuda.Event() for _ in range(n)]

        def issue_h2d(k):
            with torch.cuda.stream(copy), torch.no_grad():
                copy.wait_event(used[k])   # the step that last read this input set has finished
                for dst, src in zip(sets[k], host_in):
                    dst.copy_(src, non_blocking=True)
                h2d_done[k].record(copy)

        def run_loop(count, out_off):
            issue_h2d(0)
            for i in range(count):
                k = i % n
                main.wait_event(h2d_done[k])
                if n > 1 and i + 1 < count:
                    issue_h2d((i + 1) % n)
                flush.zero_()
                out = graphs[k].replay() if graphs[k] is not None else step(*sets[k])
                used[k].record(main)
                if n == 1 and i + 1 < count:
                    issue_h2d(0)
                host_out[out_off + i].copy_(torch.cat([out[0], out[1]]), non_blocking=True)

        for k in range(n):
            used[k].record(main)
        run_loop(warm, 0)
        barrier()
        n_alloc0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record(main)
        copy.wait_event(s_)
        run_loop(steps, warm)
        e_.record(main)
        barrier()
        t2 = torch.tensor([s_.elapsed_time(e_)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        ms = float(t2.item()) / steps
        return {"value": B_glob / (ms / 1e3), "unit": "utts/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": 8, "ms_per_step": ms, "loss": float(host_out[-1, 0]),
                "h2d": "pinned, copy stream, prefetched one step ahead (two input sets)" if n > 1 else "pinned, same stream",
                "l2_flush_inside_region": True,
                "device_allocs_in_timed_region": int(torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - n_alloc0)}

    main_res = measure(args.backward, B_local, B_global, host, want_e2e=not args.no_e2e, want_profile=True)
    wl, step, ins = main_res["wl"], main_res["step"], main_res["ins"]
    precision = main_res["precision"]
    x3 = precision in ("bf16x3", "fp16x3", "fp16m8")

    kern = main_res["kern"]
    el_d, tl_d = ins[3], ins[4]
    cells = float((el_d.double() * (tl_d.double() + 1)).sum().item())
    Vp = c["V"] + 1
    gemm_flops = 2.0 * cells * c["H"] * Vp          # one pass of the joint GEMM (algorithmic)
    issue = 2 if precision == "fp16m8" else 3 if x3 else 1   # MMA issue-equivalents per product (an e4m3 K=32 MMA = 1/2)
    roofline = None
    extra = {}
    mode_key = f"joint_fwd|{args.backward}|{precision}|{c['activation']}|B{B_local}|T{c['T']}|U{c['U']}|V{c['V']}"
    if "joint_fwd" in kern:
        ach = gemm_flops / (kern["joint_fwd"] * 1e-3) / 1e12
        traffic, tsrc = (None, None) if (c["ragged"] or c["dropout"] > 0) else traffic_from_profiles(mode_key)
        roofline = {"kernel": "joint_fwd_kernel (pass 1: joint GEMM + online log-softmax)", "bound": "tensor",
                    "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                    "traffic": traffic, "traffic_source": tsrc, "backward_mode": args.backward,
                    "peak_source": pk["source"] + " bf16 sustained",
                    "algorithmic_flops_per_launch": gemm_flops, "ms": kern["joint_fwd"],
                    "mma_issue_multiplier": issue,
                    # the same launch counted in MMA issues (what the tensor pipe actually executes)
                    "frac_of_mma_issue_rate": issue * ach / pk["tf_sustained"]}
        tensor_ms = sum(kern.get(k, 0.0) for k in ("joint_fwd", "joint_bwd_dz", "gemm_dhid", "gemm_dw"))
        ach_all = 3.0 * gemm_flops / (tensor_ms * 1e-3) / 1e12
        extra["joint_fwd_bwd"] = {"bound": "tensor", "achieved": ach_all, "peak": pk["tf_sustained"],
                                  "unit": "TFLOP/s", "frac": ach_all / pk["tf_sustained"], "ms": tensor_ms,
                                  "algorithmic_flops": 3.0 * gemm_flops,
                                  "note": "fwd + dHid + dW credited; the pass-2 logits recompute is overhead"}
        whole = 3.0 * gemm_flops / (main_res["ms"] * 1e-3) / 1e12
        extra["whole_step"] = {"bound": "tensor", "achieved": whole, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                               "frac": whole / pk["tf_sustained"], "ms": main_res["ms"]}
        if "joint_dz_sweep" in kern:
            terms = 2 if x3 else 1
            by = cells * (4.0 * ((Vp + 31) // 32 * 32) + 2.0 * terms * ((Vp + 15) // 16 * 16))
            g = by / (kern["joint_dz_sweep"] * 1e-3) / 1e9
            extra["joint_dz_sweep"] = {"bound": "hbm", "achieved": g, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                       "frac": g / pk["hbm_gbs"], "ms": kern["joint_dz_sweep"], "algorithmic_bytes": by}
        if "joint_dfg" in kern:
            by = cells * c["H"] * 4.0
            g = by / (kern["joint_dfg"] * 1e-3) / 1e9
            extra["joint_dfg"] = {"bound": "hbm", "achieved": g, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                  "frac": g / pk["hbm_gbs"], "ms": kern["joint_dfg"], "algorithmic_bytes": by}
    elif "rnnt_lse" in kern:
        # materialised mode runs the reference's sub-batch loop (fused_batch_size = 4): several launches per step
        by = 3.0 * cells * Vp * 4
        ms = kern["rnnt_lse+grad_per_step"]
        roofline = {"kernel": "rnnt_lse_gather + rnnt_grad (materialised logits)", "bound": "hbm",
                    "achieved": by / (ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": by / (ms * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": None, "ms": ms}
    if "rnnt_lattice" in kern:
        # 2 reads + 2 writes of 4 B per cell nominal; here 16 B of split probabilities read per direction
        # and 8 B written.  Latency-bound: T + U dependent steps — the floor is steps x (shuffle + ~10 dependent ops).
        by = cells * 16.0
        extra["rnnt_lattice"] = {"bound": "latency (hbm nominal)", "achieved": by / (kern["rnnt_lattice"] * 1e-3) / 1e9,
                                 "peak": pk["hbm_gbs"], "unit": "GB/s", "ms": kern["rnnt_lattice"],
                                 "algorithmic_bytes": by, "dependent_steps": c["T"] + c["U"],
                                 "ns_per_dependent_step": 1e6 * kern["rnnt_lattice"] / (c["T"] + c["U"])}
    if "ctc_lattice" in kern:
        by = 2.0 * B_local * c["T"] * Vp * 4 + 2.0 * B_local * c["T"] * (2 * c["U"] + 1) * 4
        ms = kern["ctc_lattice"] + kern.get("ctc_grad", 0.0)
        extra["ctc"] = {"bound": "latency (hbm nominal)", "achieved": by / (ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"],
                        "unit": "GB/s", "ms": ms, "algorithmic_bytes": by, "dependent_steps": c["T"]}

    # ---------------- the other backward mode, the importance-epoch step, weak scaling
    if not args.no_extras:
        other = "stash" if args.backward == "recompute" else "recompute"
        if args.mode == "tcgen05":
            r = measure(other, B_local, B_global, host)
            extra[f"{other}_mode"] = {"value": r["value"], "unit": "utts/s", "ms_per_step": r["ms"],
                                      "note": ("the forward keeps the valid cells' logits + hidden activations in HBM for the "
                                               "backward pass (RNNTJoint(backward_mode='stash'))") if other == "stash"
                                      else "logits never written to HBM"}
            del r
            r = measure(args.backward, B_local, B_global, host, variant="fisher")
            extra["fisher_accum_step"] = {"value": r["value"], "unit": "utts/s", "ms_per_step": r["ms"],
                                          "note": "importance epoch: no penalty, F += loss * grad^2 (cl_baseline_ewc.py:245-255)"}
            del r
        if world > 1 and strong:
            hw = host_inputs(B, 0, 1234 + rank)
            r = measure(args.backward, B, B * world, hw)
            extra["weak_scaling"] = {"value": r["value"], "unit": "utts/s", "ms_per_step": r["ms"], "per_gpu_batch": B,
                                     "global_batch": B * world}
            del r
        torch.cuda.empty_cache()
        if rank == 0 and world == 1:
            extra.update(standalone_probes(args, c, dev, pk))

    cpu = None
    incumbent = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = run_cpu_baseline(c, args.cpu_sample, steps=1, warmup=0)
        if args.config == 2 and not args.no_extras:
            cpu["reference_cpu_rnnt_b1"] = run_cpu_reference_rnnt(c)
    if rank == 0 and world == 1 and args.config == 2 and not args.no_extras:
        torch.cuda.empty_cache()
        incumbent = run_gpu_incumbent(c, dev)

    if rank == 0:
        metric = "RNNT+CTC+EWC fwd/bwd utts/s (B32,T250,U100,V1024)" if args.config == 2 else \
            f"RNNT+CTC+EWC fwd/bwd utts/s (B{c['B']},T{c['T']},U{c['U']},V{c['V']})"
        line = {
            "metric": metric, "value": main_res["value"], "unit": "utts/s",
            "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": main_res["ms"],
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": ("f32 (joint GEMM: fp16 hi.hi + two e4m3 correction MMAs on tcgen05, fp32 accumulate)"
                      if precision == "fp16m8" else
                      f"f32 (joint GEMM: {precision[:4]} hi/lo split x3 on tcgen05, fp32 accumulate)"
                      if x3 else "bf16 joint GEMM, fp32 elsewhere"),
            "data": "synthetic",
            "config": {"workload": c["workload"] + (", ragged lengths" if c["ragged"] and args.config == 2 else ""),
                       "global_batch": B_global, "per_gpu_batch": B_local, "activation": c["activation"],
                       "joint_dropout": c["dropout"], "joint_impl": args.mode, "precision": precision,
                       "backward_mode": args.backward + (" (logits never written to HBM)" if args.backward == "recompute" else ""),
                       "cuda_graph": main_res.get("gs") is not None, "cuda_graph_error": main_res["graph_error"],
                       "trainable_params": main_res["n_params"],
                       "l2": "256 MB buffer written between timed iterations (L2 flush)",
                       "parallelism": f"dp{world}: global batch sharded {B_global}/{world} per GPU, one flat-gradient NCCL "
                                      f"all-reduce per step" if strong else f"dp{world}: {B} utterances per GPU (weak)"},
            "e2e": main_res.get("e2e"),
            "gpu_launches": main_res["launches_per_step"] * steps, "gpu_launches_per_step": main_res["launches_per_step"],
            "clocks": main_res["clocks"], "roofline": roofline, "roofline_more": extra, "kernel_ms": kern,
            "cpu_baseline": cpu, "gpu_incumbent": incumbent, "impl": "b200",
        }
        print(json.dumps(line), flush=True)
    finish(world)


def finish(world):
    """Leave without tearing the NCCL communicator down: destroy_process_group() with CUDA graphs that captured
    collectives still alive can block for the watchdog timeout.  Every rank has finished its work when it gets here."""
    if world > 1:
        import torch.distributed as dist

        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def run_gpu_incumbent(c, dev, steps=3):
    """The reference's own GPU path on the same B200 (the like-for-like bar): torch joint
    in sub-batches of fused_batch_size = 4 (modules/rnnt.py:1425: cuBLAS GEMMs, materialised hidden + logits), the
    reference's numba-CUDA RNNTLossNumba (staged byte-for-byte under oracle/_ref by oracle/stage_ref.py), ATen
    torch.nn.CTCLoss, and the per-tensor EWC Python loop with its .item() (cl_baseline_ewc.py:69-81).  Returns a dict,
    or {"unavailable": why}."""
    try:
        from oracle import stage_ref
        if not stage_ref.available():
            return {"unavailable": "oracle/_ref not staged"}
        from numba import cuda as ncuda
        if not ncuda.is_available():
            return {"unavailable": "numba.cuda not available on this box"}
        RNNTLossNumba = stage_ref.load_rnnt_loss_numba()
        torch.manual_seed(1234)
        B, V = c["B"], c["V"]
        enc_l = torch.nn.Linear(c["D_enc"], c["H"]).to(dev)
        pred_l = torch.nn.Linear(c["D_pred"], c["H"]).to(dev)
        out_l = torch.nn.Linear(c["H"], V + 1).to(dev)
        head = torch.nn.Conv1d(c["D_enc"], V + 1, kernel_size=1).to(dev)
        params = {n: p_ for m_, pre in ((enc_l, "enc."), (pred_l, "pred."), (out_l, "out."), (head, "ctc."))
                  for n, p_ in ((pre + k, v) for k, v in m_.named_parameters())}
        star = {k: v.detach() + 0.01 * torch.randn_like(v) for k, v in params.items()}
        fish = {k: torch.rand_like(v) for k, v in params.items()}
        act = {"tanh": torch.tanh, "relu": torch.relu, "sigmoid": torch.sigmoid}[c["activation"]]
        loss_fn = RNNTLossNumba(blank=V, reduction="none")
        ctc_fn = torch.nn.CTCLoss(blank=V, reduction="none", zero_infinity=True)
        enc, dec, tr, el, tl = [x.to(dev) for x in synth(dict(c, ragged=False), B, 1234)]

        def step():
            for p_ in params.values():
                p_.grad = None
            e_ = enc.clone().requires_grad_(True)
            d_ = dec.clone().requires_grad_(True)
            f = enc_l(e_.transpose(1, 2))
            g = pred_l(d_.transpose(1, 2))
            losses = []
            for b0 in range(0, B, 4):   # the reference's fused sub-batch loop
                sl = slice(b0, b0 + 4)
                z = out_l(act(f[sl].unsqueeze(2) + g[sl].unsqueeze(1)))
                losses.append(loss_fn(z, tr[sl].contiguous(), el[sl], tl[sl]))
            l_rnnt = torch.cat(losses).mean()
            lp = head(e_).transpose(1, 2).log_softmax(-1)
            l_ctc = ctc_fn(lp.transpose(0, 1), tr, el, tl).mean()
            loss = (1 - CTC_WEIGHT) * l_rnnt + CTC_WEIGHT * l_ctc
            pen_avg = 0.0
            for k, p_ in params.items():   # get_penalty_grads + set_grads, per tensor, one host sync
                p_.grad = E_LAMBDA * 2 * fish[k] * (p_.data - star[k])
                pen_avg = pen_avg + torch.mean(torch.abs(p_.grad))
            pen_avg = (pen_avg / len(params)).item()
            loss.backward()
            return loss

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            loss = step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / steps
        return {"value": B / (ms * 1e-3), "unit": "utts/s", "ms_per_step": ms, "loss": float(loss),
                "what": "reference GPU path on this B200: torch joint (fused_batch_size 4, cuBLAS fp32) + the reference's "
                        "numba-CUDA RNNTLossNumba (oracle/_ref) + ATen CTCLoss + per-tensor EWC loop"}
    except Exception as ex:   # numba / NVVM mismatches must not take the bench line down
        return {"unavailable": f"{type(ex).__name__}: {ex}"[:300]}


def run_cpu_reference_rnnt(c):
    """The reference's OWN CPU transducer loss (un-jitted Python, cpu_rnnt.py:164-422 through RNNTLossNumba on CPU
    tensors, staged under oracle/_ref) on ONE utterance of the named shape with a reduced vocabulary slice for the
    log-softmax input — the loop the C port in cpu_baseline replaces.  ~20 s; strictly per-sample (cpu_rnnt.py:365)."""
    try:
        from oracle import stage_ref
        if not stage_ref.available():
            return {"unavailable": "oracle/_ref not staged"}
        RNNTLossNumba = stage_ref.load_rnnt_loss_numba()
        torch.manual_seed(0)
        T, U, Vp = c["T"], c["U"], c["V"] + 1
        z = torch.randn(1, T, U + 1, Vp, requires_grad=True)
        lab = torch.randint(0, Vp - 1, (1, U))
        t0 = time.perf_counter()
        loss = RNNTLossNumba(blank=Vp - 1, reduction="sum")(z, lab, torch.tensor([T]), torch.tensor([U]))
        loss.backward()
        dt = time.perf_counter() - t0
        return {"value": 1.0 / dt, "unit": "utts/s (transducer loss fwd+bwd only)", "seconds": dt, "kind": "reference",
                "cores": 1, "sample": f"1 utterance, T={T} U={U} V={Vp - 1}: RNNTLossNumba on CPU tensors (log_softmax + "
                                      "the pure-Python alpha/beta/grad loops)"}
    except Exception as ex:
        return {"unavailable": f"{type(ex).__name__}: {ex}"[:300]}


def _time_ms(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def standalone_probes(args, c, dev, pk):
    """HBM-roofline probes of the kernels the step does not exercise at a size that means anything: the drop-in
    transducer loss on materialised logits, CTC at config-5 size, and every regulariser sweep at config-4 size
    (120 M parameters: working sets far larger than L2)."""
    from indic_cl_asr_b200 import CTCLoss, RNNTLossNumba, _lib, cl
    from indic_cl_asr_b200.cl.flat import FlatDict, Layout

    out = {}
    L = _lib.lib()

    def hbm(by, ms, **kw):
        g = by / (ms * 1e-3) / 1e9
        return dict({"bound": "hbm", "achieved": g, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": g / pk["hbm_gbs"],
                     "ms": ms, "algorithmic_bytes": by}, **kw)

    # ---- drop-in RNNTLoss on materialised logits (a5/a6): one sub-batch of 4 like the reference's fused loop
    b, T, U1, Vp = 4, 250, 101, 1025
    z = torch.randn(b, T, U1, Vp, device=dev, requires_grad=True)
    lab = torch.randint(0, Vp - 1, (b, U1 - 1), device=dev)
    al = torch.full((b,), T, dtype=torch.long, device=dev)
    ll = torch.full((b,), U1 - 1, dtype=torch.long, device=dev)
    loss = RNNTLossNumba(blank=Vp - 1, reduction="sum")

    def rnnt_step():
        z.grad = None
        loss(z, lab, al, ll).backward()

    rnnt_step()
    L.clasr_set_profiling(1)
    L.clasr_profile_reset()
    for _ in range(5):
        rnnt_step()
    torch.cuda.synchronize()
    lse, grad = _lib.profile_ms("rnnt_lse"), _lib.profile_ms("rnnt_grad")
    L.clasr_set_profiling(0)
    cells = b * T * U1
    out["rnnt_materialised"] = hbm(3.0 * cells * Vp * 4, lse + grad, lse_ms=lse, grad_ms=grad,
                                   note="drop-in RNNTLoss, [4,250,101,1025] logits (1.24 GB read twice, written once)")
    del z
    # ---- CTC forward-backward at config-5 size
    B5, T5, V5, U5 = 64, 500, 4096, 200
    lp = torch.randn(B5, T5, V5 + 1, device=dev).log_softmax(-1).requires_grad_(True)
    tg = torch.randint(0, V5, (B5, U5), device=dev)
    il = torch.full((B5,), T5, dtype=torch.long, device=dev)
    tl = torch.full((B5,), U5, dtype=torch.long, device=dev)
    ctc = CTCLoss(num_classes=V5, zero_infinity=True)

    def ctc_step():
        lp.grad = None
        ctc(log_probs=lp, targets=tg, input_lengths=il, target_lengths=tl).backward()

    ctc_step()
    L.clasr_set_profiling(1)
    L.clasr_profile_reset()
    for _ in range(5):
        ctc_step()
    torch.cuda.synchronize()
    lat, gr = _lib.profile_ms("ctc_lattice"), _lib.profile_ms("ctc_grad")
    L.clasr_set_profiling(0)
    by = 2.0 * B5 * T5 * (V5 + 1) * 4 + 2.0 * B5 * T5 * (2 * U5 + 1) * 4
    out["ctc_config5"] = hbm(by, lat + gr, lattice_ms=lat, grad_ms=gr,
                             note="B=64 T=500 V=4096 U=200; the lattice part is 500 dependent steps")
    del lp
    # ---- regulariser sweeps over a flat 120 M-parameter buffer (config 4)
    P = args.ewc_params
    if P > 0:
        lay = Layout([("flat", torch.Size([P]))])
        th, st_, fi = [FlatDict(lay, torch.randn(lay.total, device=dev)) for _ in range(3)]
        g = torch.randn(lay.total, device=dev)
        o = torch.empty(lay.total, device=dev)
        w = torch.ones(1, device=dev)
        val = torch.zeros(1, dtype=torch.float64, device=dev)
        s = lambda: _lib.stream_ptr(dev)
        ms = _time_ms(lambda: cl.get_penalty_grads_async(E_LAMBDA, fi, th, st_, out=o))
        out["ewc_penalty_grad_120M"] = hbm(16.0 * P, ms, note="includes the tiny penalty_avg kernel")
        ms = _time_ms(lambda: L.clasr_cl_fisher_accum(fi.flat.data_ptr(), g.data_ptr(), lay.total, w.data_ptr(), s()))
        out["fisher_accum_120M"] = hbm(12.0 * P, ms)
        ms = _time_ms(lambda: L.clasr_cl_mas_accum(fi.flat.data_ptr(), g.data_ptr(), lay.total, s()))
        out["mas_accum_120M"] = hbm(12.0 * P, ms)
        ms = _time_ms(lambda: L.clasr_cl_penalty_value_grad(th.flat.data_ptr(), st_.flat.data_ptr(), fi.flat.data_ptr(),
                                                            lay.total, 1.0, val.data_ptr(), g.data_ptr(), s()))
        out["mas_penalty_value_grad_120M"] = hbm(20.0 * P, ms, note="3 reads + gradient read-modify-write")
        ms = _time_ms(lambda: L.clasr_cl_scale_merge(th.flat.data_ptr(), fi.flat.data_ptr(), lay.total, 100.0, 0.9, 0, s()))
        out["fisher_merge_120M"] = hbm(16.0 * P, ms, note="F /= n; F_main = gamma F_main + F (2 reads, 2 writes)")
        ms = _time_ms(lambda: L.clasr_cl_snapshot(o.data_ptr(), th.flat.data_ptr(), lay.total, s()))
        out["theta_snapshot_120M"] = hbm(8.0 * P, ms)
    return out


def main_config4(args, dev, world, rank):
    """configs[3]: MAS importance accumulation + penalty over a 120 M-parameter flat buffer after a task switch, and
    the per-task NCCL all-reduce of the importance buffer.  A 'step' = |grad| accumulation sweep + penalty value and
    gradient sweep; the all-reduce (once per task, 480 MB) is timed separately and reported as bus bandwidth."""
    import torch.distributed as dist

    from indic_cl_asr_b200 import _lib
    from indic_cl_asr_b200.dist import allreduce_importance_

    pk = peaks()
    L = _lib.lib()
    P = (args.ewc_params + 3) // 4 * 4
    theta, star, omega, grad = [torch.randn(P, device=dev) for _ in range(4)]
    omega.abs_()
    val = torch.zeros(1, dtype=torch.float64, device=dev)
    s = lambda: _lib.stream_ptr(dev)

    def step():
        L.clasr_cl_mas_accum(omega.data_ptr(), grad.data_ptr(), P, s())                      # cl_baseline_mas.py:267-270
        L.clasr_cl_penalty_value_grad(theta.data_ptr(), star.data_ptr(), omega.data_ptr(), P, 1.0, val.data_ptr(),
                                      grad.data_ptr(), s())                                  # :70-75, 231-234

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(int(os.environ.get("LOCAL_RANK", "0"))) as clk:
        barrier()
        for a, b in evs:
            a.record()
            step()
            b.record()
        barrier()
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps
    by = 32.0 * P    # accumulate: 2 reads + 1 write; penalty: 3 reads + grad read-modify-write
    ar = None
    if world > 1:
        cnt = torch.ones(1, device=dev)
        for _ in range(2):
            allreduce_importance_(omega, cnt)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            allreduce_importance_(omega, cnt)
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b) / 5], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ar_ms = float(t.item())
        ar = {"ms": ar_ms, "bytes": 4.0 * P, "algbw_gbs": 4.0 * P / (ar_ms * 1e-3) / 1e9,
              "busbw_gbs": 4.0 * P / (ar_ms * 1e-3) / 1e9 * 2 * (world - 1) / world,
              "nvlink_peak_gbs_per_direction": 900.0}
    if rank == 0:
        g = by / (ms * 1e-3) / 1e9
        print(json.dumps({
            "metric": "MAS importance accumulation + penalty sweeps per second (120M-param flat buffer)",
            "value": 1e3 / ms, "unit": "sweeps/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "replicas only (state is replicated; one all-reduce per task)",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[3]: MAS importance accumulation + penalty over a 120M-param flat buffer, "
                                   "NCCL all-reduce of the importance buffer once per task", "params": P,
                       "l2": "working set 1.9 GB per sweep, far larger than L2"},
            "roofline": {"bound": "hbm", "achieved": g, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": g / pk["hbm_gbs"],
                         "traffic": None, "algorithmic_bytes": by},
            "importance_allreduce": ar, "clocks": clk.summary(), "gpu_launches": 2 * args.steps, "impl": "b200"}),
            flush=True)
    finish(world)


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    c = resolve_cfg(args)
    r = run_cpu_baseline(c, args.cpu_sample, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "RNNT+CTC+EWC fwd/bwd utts/s (B32,T250,U100,V1024)", "value": r["value"],
        "unit": "utts/s", "n_gpus": world, "steps": args.steps, "warmup": min(args.warmup, 1),
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": c["workload"], "sample": r["sample"]},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "utts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_b200(a)
