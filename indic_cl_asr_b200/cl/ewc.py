"""EWC hooks (reference cl_baseline_ewc.py) as single fused sweeps over flat buffers."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from .. import _lib
from .flat import FlatDict, as_flat

__all__ = ["get_penalty_grads", "get_penalty_grads_async", "fisher_accumulate", "fisher_finalise"]


def _e_lambda(config) -> float:
    if hasattr(config, "cl_config"):
        return float(config.cl_config.e_lambda)
    if isinstance(config, dict):
        return float(config["cl_config"]["e_lambda"])
    return float(config)


def get_penalty_grads_async(config, fish, curr_checkpoint, checkpoint, out: Optional[torch.Tensor] = None,
                            accumulate: bool = False) -> Tuple[FlatDict, torch.Tensor]:
    """As get_penalty_grads but returns ``penalty_avg`` as a 1-element device tensor (no host sync).

    ``out``: flat fp32 buffer to write (or, with ``accumulate``, add) the penalty gradient into — pass the model's
    flat gradient buffer to make ``set_grads`` a no-copy re-pointing."""
    theta = as_flat(curr_checkpoint)
    layout = theta.layout
    star = as_flat(checkpoint, layout)
    F = as_flat(fish, layout)
    dev = theta.flat.device
    if out is None:
        out = torch.empty(layout.total, dtype=torch.float32, device=dev)
        if accumulate:
            out.zero_()
    items, seg_numel, n_items = layout.device_tables(dev)
    n_seg = len(layout.names)
    seg_abs = torch.zeros(n_seg, dtype=torch.float64, device=dev)
    avg = torch.empty(1, dtype=torch.float32, device=dev)
    L = _lib.lib()
    coef = _e_lambda(config) * 2  # cl_baseline_ewc.py:74: e_lambda * 2 * F * (theta - theta*)
    with torch.cuda.device(dev):
        s = _lib.stream_ptr(dev)
        _lib.check(L.clasr_cl_penalty_grad(theta.flat.data_ptr(), star.flat.data_ptr(), F.flat.data_ptr(),
                                           out.data_ptr(), items.data_ptr(), n_items, coef, int(accumulate),
                                           seg_abs.data_ptr(), s), "cl_penalty_grad")
        _lib.check(L.clasr_cl_penalty_avg(seg_abs.data_ptr(), seg_numel.data_ptr(), n_seg, avg.data_ptr(), s),
                   "cl_penalty_avg")
    return FlatDict(layout, out), avg


def get_penalty_grads(config, fish, curr_checkpoint, checkpoint) -> Tuple[Dict[str, torch.Tensor], float]:
    """Reference signature and return (cl_baseline_ewc.py:69-81): ``(result dict, penalty_avg float)``.
    The float forces the same single host sync per step the reference has (``.item()``, :81)."""
    result, avg = get_penalty_grads_async(config, fish, curr_checkpoint, checkpoint)
    return result, avg.item()


def fisher_accumulate(fish, curr_grads, loss: torch.Tensor) -> None:
    """fish[key] += mean(loss) * grad[key]**2 for every key with a gradient (cl_baseline_ewc.py:245-255).
    ``fish`` must be flat (get_zero_params); ``curr_grads`` may be a FlatDict (the flat gradient buffer) or
    the plain dict get_grads() returns.  The loss value is read on the device."""
    if not isinstance(fish, FlatDict) or not fish.is_intact():
        raise TypeError("fisher_accumulate: `fish` must come from get_zero_params()")
    layout = fish.layout
    if isinstance(curr_grads, FlatDict) and curr_grads.is_intact() and curr_grads.layout.key() == layout.key():
        g = curr_grads.flat
    else:
        g = torch.zeros(layout.total, dtype=torch.float32, device=fish.flat.device)
        gv = layout.views(g)
        for k, v in curr_grads.items():
            if k in gv:  # keys outside the trainable set would raise KeyError in the reference's fish[key]
                gv[k].copy_(v)
            else:
                raise KeyError(k)
    w = torch.mean(loss.detach().clone()).to(torch.float32).reshape(1)
    dev = fish.flat.device
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().clasr_cl_fisher_accum(fish.flat.data_ptr(), g.data_ptr(), layout.total, w.data_ptr(),
                                                    _lib.stream_ptr(dev)), "cl_fisher_accum")


def fisher_finalise(fish: FlatDict, main_fish: Optional[FlatDict], total_ds, e_gamma: float) -> FlatDict:
    """fish /= total_ds; main = fish (first task) else gamma*main + fish (cl_baseline_ewc.py:267-280)."""
    dev = fish.flat.device
    L = _lib.lib()
    with torch.cuda.device(dev):
        s = _lib.stream_ptr(dev)
        if main_fish is None:
            _lib.check(L.clasr_cl_scale_merge(fish.flat.data_ptr(), fish.flat.data_ptr(), fish.layout.total,
                                              float(total_ds), 0.0, 1, s), "cl_scale_merge")
            return fish
        main = as_flat(main_fish, fish.layout)
        _lib.check(L.clasr_cl_scale_merge(main.flat.data_ptr(), fish.flat.data_ptr(), fish.layout.total,
                                          float(total_ds), float(e_gamma), 0, s), "cl_scale_merge")
        return main
