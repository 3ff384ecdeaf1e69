#!/usr/bin/env python
"""bench.py — RNNT+CTC+EWC fwd/bwd utterances/s at B=32, T=250, U=100, V=1024 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port), rank 0 only

One "step" = one pass of the hot path over one batch of synthetic input:
  fused joint (enc/pred projections -> tcgen05 joint GEMM + online log-softmax) -> alpha/beta wavefront -> RNNT cost,
  CTC head (Conv1d k=1 -> log_softmax kernel) -> CTC forward-backward, mixed loss 0.7/0.3, backward to d enc_out,
  d pred_out and every joint / CTC-head parameter, EWC penalty gradient pre-loaded into the flat gradient buffer
  (cl_baseline_ewc.py:228-240), and - for N > 1 - ONE NCCL sum all-reduce of the flat gradient buffer.
Scaling is WEAK: every rank processes its own B=32 utterances (global batch 32*N).
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

CFG = dict(B=32, T=250, U=100, V=1024, H=640, D_enc=512, D_pred=640)
CTC_WEIGHT = 0.3   # cfg.aux_ctc.ctc_loss_weight (hybrid_rnnt_ctc_models.py:233,902)
E_LAMBDA = 10.0    # config.yaml e_lambda


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="tcgen05", choices=["tcgen05", "materialised"])
    ap.add_argument("--precision", default="auto", choices=["auto", "bf16x3", "fp16x3", "bf16"],
                    help="auto = the module default (fp16x3)")
    ap.add_argument("--activation", default="tanh", choices=["tanh", "relu", "sigmoid"])
    ap.add_argument("--ragged", type=int, default=0)
    ap.add_argument("--dropout", type=float, default=0.0, help="joint dropout (the shipped checkpoint trains with 0.2)")
    ap.add_argument("--cpu-sample", type=int, default=4, help="utterances in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-overlap-ctc", action="store_true", help="keep the CTC branch on the main stream")
    ap.add_argument("--ewc-params", type=int, default=120_000_000, help="size of the stand-alone regulariser sweep probe")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def synth(B, device, seed, ragged):
    g = torch.Generator().manual_seed(seed)
    c = CFG
    enc = torch.randn(B, c["D_enc"], c["T"], generator=g)
    dec = torch.randn(B, c["D_pred"], c["U"] + 1, generator=g)
    tr = torch.randint(0, c["V"], (B, c["U"]), generator=g)
    if ragged:
        el = torch.randint(c["T"] // 2, c["T"] + 1, (B,), generator=g)
        tl = torch.randint(c["U"] // 2, c["U"] + 1, (B,), generator=g)
        el[0], tl[0] = c["T"], c["U"]
    else:
        el, tl = torch.full((B,), c["T"]), torch.full((B,), c["U"])
    return enc, dec, tr, el, tl


# ----------------------------------------------------------------------------------------------- CPU baseline
def cpu_reference_step(sample, activation, modules_state=None):
    """The reference's CPU path on a bounded sample: restated joint (torch CPU, all threads) + transducer loss
    (oracle/lattice.c, a C port of cpu_rnnt.py) + torch.nn.CTCLoss + dict-of-tensors EWC (oracle/cl_oracle.py)."""
    from oracle import c_port, cl_oracle, joint_oracle

    c = CFG
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
                                                              "out.weight", "out.bias")}, activation)
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


def run_cpu_baseline(n_utts, activation, steps=1, warmup=0):
    torch.set_num_threads(os.cpu_count() or 1)
    sample = synth(n_utts, "cpu", 1234, 0)
    state = None
    for _ in range(warmup):
        _, _, state = cpu_reference_step(sample, activation, state)
    times = []
    for _ in range(steps):
        dt, loss, state = cpu_reference_step(sample, activation, state)
        times.append(dt)
    total = sum(times)
    return dict(value=n_utts * steps / total, unit="utts/s", cores=torch.get_num_threads(), kind="port",
                sample=f"{n_utts} utterances of the same shape (T={CFG['T']},U={CFG['U']},V={CFG['V']},H={CFG['H']}), "
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
            self._stop.wait(0.01 if self._nvml else 0.2)

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
def main_b200(args):
    import torch.distributed as dist

    from indic_cl_asr_b200 import CTCLoss, ConvASRDecoder, HybridRNNTCTCLoss, RNNTJoint, RNNTLoss, _lib, cl
    from indic_cl_asr_b200.dist import allreduce_flat_

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    torch.backends.cudnn.allow_tf32 = False       # "fp32" config: torch's own small GEMMs/convs stay true fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    c = CFG
    pk = peaks()
    L = _lib.lib()

    torch.manual_seed(1234)
    joint = RNNTJoint(jointnet=dict(encoder_hidden=c["D_enc"], pred_hidden=c["D_pred"], joint_hidden=c["H"],
                                    activation=args.activation, dropout=args.dropout),
                      num_classes=c["V"], fuse_loss_wer=True, fused_batch_size=4, fused_impl=args.mode,
                      precision=args.precision).to(dev)
    args.precision = joint.precision   # 'auto' resolved by the module
    joint.set_loss(RNNTLoss(num_classes=c["V"]))
    joint.set_wer(object())
    head = ConvASRDecoder(feat_in=c["D_enc"], num_classes=c["V"]).to(dev)
    ctc = CTCLoss(num_classes=c["V"], zero_infinity=True)
    model = torch.nn.ModuleDict({"joint": joint, "ctc_decoder": head})
    hybrid = HybridRNNTCTCLoss(joint, head, ctc, ctc_loss_weight=CTC_WEIGHT, overlap_ctc=not args.no_overlap_ctc)
    fp = cl.flat_params(model)
    theta = cl.get_params(model)
    star = cl.get_params_clone(model)
    star.flat.add_(0.01 * torch.randn_like(star.flat))
    fish = cl.get_zero_params(model, dev)
    fish.flat.uniform_(0.0, 1.0)
    ewc_cfg = {"cl_config": {"e_lambda": E_LAMBDA}}

    enc_h, dec_h, tr_h, el_h, tl_h = [x.pin_memory() for x in synth(c["B"], "cpu", 1234 + rank, args.ragged)]
    enc_d, dec_d = enc_h.to(dev).requires_grad_(True), dec_h.to(dev).requires_grad_(True)
    tr_d, el_d, tl_d = tr_h.to(dev), el_h.to(dev), tl_h.to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step(enc, dec, tr, el, tl):
        fp.bind_grads(zero=False)
        pen, avg = cl.get_penalty_grads_async(ewc_cfg, fish, theta, star, out=fp.grad)  # writes (=) the penalty grads
        enc.grad = None
        dec.grad = None
        loss, _ = hybrid(enc, el, dec, tr, tl)   # training_step's loss half (hybrid_rnnt_ctc_models.py:868-902)
        (loss / world).backward()            # local mean_batch / N: the SUM all-reduce gives the global mean's gradient
        if world > 1:
            allreduce_flat_(fp.grad)
        return loss, avg

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(enc_d, dec_d, tr_d, el_d, tl_d)
    barrier()

    # ---------------- device-resident timing (value)
    n0 = _lib.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local_rank) as clk:
        barrier()
        for s_, e_ in evs:
            flush.zero_()                      # evict L2 between timed iterations (outside the event pair)
            s_.record()
            step(enc_d, dec_d, tr_d, el_d, tl_d)
            e_.record()
        barrier()
    total_ms = sum(s_.elapsed_time(e_) for s_, e_ in evs)
    launches = (_lib.launch_count() - n0) // max(1, args.steps)
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = c["B"] * world * args.steps / (total_ms / 1e3)

    # ---------------- end-to-end through the public API with HOST inputs (pinned H2D + D2H of the loss every step)
    e2e = None
    if not args.no_e2e:
        h2d = sum(x.numel() * x.element_size() for x in (enc_h, dec_h, tr_h, el_h, tl_h))
        host_out = torch.empty(2, dtype=torch.float32).pin_memory()
        evs2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        n_warm = max(args.warmup, 3) + 3
        n_dev_alloc0 = 0
        # ONE loop for warm-up and timed steps: identical tensor lifetimes, so the caching allocator (incl. the CTC
        # side-stream pool) is in steady state when the timed region starts — a cudaMalloc inside it stalls the queue
        for it in range(n_warm + args.steps):
            timed = it >= n_warm
            if it == n_warm:
                barrier()
                n_dev_alloc0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
            flush.zero_()
            if timed:
                evs2[it - n_warm][0].record()
            e1 = enc_h.to(dev, non_blocking=True).requires_grad_(True)
            d1 = dec_h.to(dev, non_blocking=True).requires_grad_(True)
            loss, avg = step(e1, d1, tr_h.to(dev, non_blocking=True), el_h.to(dev, non_blocking=True),
                             tl_h.to(dev, non_blocking=True))
            host_out.copy_(torch.stack([loss.detach().float().reshape(()), avg.reshape(())]), non_blocking=True)
            if timed:
                evs2[it - n_warm][1].record()
        barrier()
        t2 = torch.tensor([sum(s_.elapsed_time(e_) for s_, e_ in evs2)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e = {"value": c["B"] * world * args.steps / (float(t2.item()) / 1e3), "unit": "utts/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 8, "ms_per_step": float(t2.item()) / args.steps,
               "loss": float(host_out[0]),
               # cudaMalloc calls inside the timed region (0 = the caching allocator is in steady state)
               "device_allocs_in_timed_region": int(torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - n_dev_alloc0)}

    # ---------------- per-kernel durations (library-side CUDA events on the launch stream) -> roofline
    hybrid.overlap_ctc = False   # per-kernel durations must not include time shared with the side-stream CTC branch
    L.clasr_set_profiling(1)
    L.clasr_profile_reset()
    for _ in range(min(args.steps, 5)):
        flush.zero_()
        step(enc_d, dec_d, tr_d, el_d, tl_d)
    torch.cuda.synchronize()
    kern = {k: _lib.profile_ms(k) for k in ("joint_fwd", "joint_bwd_dz", "joint_dz_sweep", "gemm_dhid", "gemm_dw",
                                            "joint_dfg", "rnnt_lattice", "rnnt_lse", "rnnt_grad", "ctc_lattice", "ctc_grad",
                                            "cl_penalty_grad")}
    kern = {k: v for k, v in kern.items() if v >= 0}
    L.clasr_set_profiling(0)
    cells = float((el_d.double() * (tl_d.double() + 1)).sum().item())
    gemm_flops = 2.0 * cells * c["H"] * (c["V"] + 1)          # one pass of the joint GEMM (algorithmic)
    roofline = None
    extra_roof = {}
    if "joint_fwd" in kern:
        ach = gemm_flops / (kern["joint_fwd"] * 1e-3) / 1e12
        roofline = {"kernel": "joint_fwd_kernel (pass 1: joint GEMM + online log-softmax)", "bound": "tensor",
                    "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                    # dram__bytes_read.sum + dram__bytes_write.sum of one launch, `ncu --set full`.
                    # stash mode (profiles/r01k_ncu_stash_mode.md): 0.032 GB read + 5.460 GB written — the algorithmic
                    # output is the kept logits and hidden activations, 4*pad32(V+1) + 4*H bytes per cell = 5.48 GB.
                    # recompute mode (r01bcd, r01g capture): 31.5 MB read, ~0 written (outputs 20 B per cell).
                    "traffic": (None if (args.precision not in ("bf16x3", "fp16x3") or args.ragged or args.dropout > 0)
                                else 31.5e6 if os.environ.get("CLASR_JOINT_STASH", "") == "0" else 5.492e9),
                    "backward_mode": "recompute" if os.environ.get("CLASR_JOINT_STASH", "") == "0" else "stash",
                    "peak_source": pk["source"] + " bf16 sustained",
                    "algorithmic_flops_per_launch": gemm_flops, "ms": kern["joint_fwd"],
                    "mma_issue_multiplier": 3 if args.precision in ("bf16x3", "fp16x3") else 1,
                    # the same launch counted in MMA issues (what the tensor pipe actually executes)
                    "frac_of_mma_issue_rate": (3 if args.precision in ("bf16x3", "fp16x3") else 1) * ach / pk["tf_sustained"]}
        tensor_ms = sum(kern.get(k, 0.0) for k in ("joint_fwd", "joint_bwd_dz", "gemm_dhid", "gemm_dw"))
        ach_all = 3.0 * gemm_flops / (tensor_ms * 1e-3) / 1e12
        extra_roof["joint_fwd_bwd"] = {"bound": "tensor", "achieved": ach_all, "peak": pk["tf_sustained"],
                                       "unit": "TFLOP/s", "frac": ach_all / pk["tf_sustained"], "ms": tensor_ms,
                                       "algorithmic_flops": 3.0 * gemm_flops,
                                       "note": "fwd + dHid + dW credited; a pass-2 logits recompute "
                                               "(CLASR_JOINT_STASH=0) is overhead"}
        if "joint_dz_sweep" in kern:
            # dZ from the kept logits: reads z fp32 [cells, round_up(Vp,32)], writes dZ bf16 hi+lo [cells, round_up(Vp,16)]
            vp = c["V"] + 1
            terms = 2 if args.precision in ("bf16x3", "fp16x3") else 1
            by = cells * (4.0 * ((vp + 31) // 32 * 32) + 2.0 * terms * ((vp + 15) // 16 * 16))
            ach_dz = by / (kern["joint_dz_sweep"] * 1e-3) / 1e9
            extra_roof["joint_dz_sweep"] = {"bound": "hbm", "achieved": ach_dz, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                            "frac": ach_dz / pk["hbm_gbs"], "ms": kern["joint_dz_sweep"],
                                            "algorithmic_bytes": by}
    elif "rnnt_lse" in kern:
        # materialised mode runs the reference's sub-batch loop (fused_batch_size = 4): several launches per step
        by = 3.0 * cells * (c["V"] + 1) * 4
        n_steps = min(args.steps, 5)
        ms = sum(m * n for m, n in (_lib.profile_ms_count(k) for k in ("rnnt_lse", "rnnt_grad")) if n > 0) / n_steps
        roofline = {"kernel": "rnnt_lse_gather + rnnt_grad (materialised logits)", "bound": "hbm",
                    "achieved": by / (ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": by / (ms * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": None, "ms": ms}

    # ---------------- stand-alone regulariser sweep (config 4 size) against the HBM roofline
    if rank == 0 and args.ewc_params > 0:
        P = args.ewc_params
        from indic_cl_asr_b200.cl.flat import FlatDict, Layout
        lay = Layout([("flat", torch.Size([P]))])
        th, st_, fi = [FlatDict(lay, torch.randn(lay.total, device=dev)) for _ in range(3)]
        out = torch.empty(lay.total, device=dev)
        for _ in range(3):
            cl.get_penalty_grads_async(E_LAMBDA, fi, th, st_, out=out)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            cl.get_penalty_grads_async(E_LAMBDA, fi, th, st_, out=out)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = statistics.median(ts)
        gbs = 16.0 * P / (ms * 1e-3) / 1e9
        extra_roof["ewc_penalty_grad_120M"] = {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                               "frac": gbs / pk["hbm_gbs"], "ms": ms, "algorithmic_bytes": 16.0 * P,
                                               "note": "includes the tiny penalty_avg kernel; 12 GB working set > L2"}
        del th, st_, fi, out

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = run_cpu_baseline(args.cpu_sample, args.activation, steps=1, warmup=0)

    if rank == 0:
        line = {
            "metric": "RNNT+CTC+EWC fwd/bwd utts/s (B32,T250,U100,V1024)", "value": value, "unit": "utts/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": (f"f32 (joint GEMM: {args.precision[:4]} hi/lo split x3 on tcgen05, fp32 accumulate)"
                      if args.precision in ("bf16x3", "fp16x3") else "bf16 joint GEMM, fp32 elsewhere"),
            "data": "synthetic",
            "config": {"workload": "configs[1]: standalone RNNT+CTC(+EWC) loss fwd/bwd, B=32 T=250 U=100 V=1024 H=640 "
                                   "per GPU, full-length utterances" + (" (ragged)" if args.ragged else ""),
                       "per_gpu_batch": c["B"], "global_batch": c["B"] * world, "activation": args.activation, "joint_dropout": args.dropout,
                       "joint_impl": args.mode, "precision": args.precision,
                       "l2": "256 MB buffer written between timed iterations (L2 flush)",
                       "parallelism": f"dp{world}: batch-sharded, one flat-gradient NCCL all-reduce per step"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clk.summary(), "roofline": roofline,
            "roofline_more": extra_roof, "kernel_ms": kern, "cpu_baseline": cpu, "impl": "b200",
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    r = run_cpu_baseline(args.cpu_sample, args.activation, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "RNNT+CTC+EWC fwd/bwd utts/s (B32,T250,U100,V1024)", "value": r["value"],
        "unit": "utts/s", "n_gpus": world, "steps": args.steps, "warmup": min(args.warmup, 1),
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: standalone RNNT+CTC(+EWC) loss fwd/bwd, B=32 T=250 U=100 V=1024 H=640",
                   "sample": r["sample"]},
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
