"""MAS hooks (reference cl_baseline_mas.py) as single fused sweeps over flat buffers."""
from __future__ import annotations

import torch

from .. import _lib
from .flat import FlatDict, as_flat, flat_params

__all__ = ["penalty", "penalty_into_grads", "mas_accumulate", "mas_finalise"]


class _MasPenaltyFn(torch.autograd.Function):
    """value = sum Omega (theta - theta*)^2 over the flat buffer; backward = 2 Omega (theta - theta*) * grad_out,
    returned as views of ONE flat buffer (one kernel instead of an autograd graph over every parameter)."""

    @staticmethod
    def forward(ctx, fp, omega_flat, star_flat, *params):
        dev = fp.device
        value = torch.zeros(1, dtype=torch.float64, device=dev)
        fp.ensure_theta_views()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().clasr_cl_penalty_value_grad(
                fp.theta.data_ptr(), star_flat.data_ptr(), omega_flat.data_ptr(), fp.layout.total, 0.0,
                value.data_ptr(), 0, _lib.stream_ptr(dev)), "cl_penalty_value_grad")
        ctx.fp, ctx.omega, ctx.star = fp, omega_flat, star_flat
        return value.to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        fp = ctx.fp
        dev = fp.device
        g = torch.zeros(fp.layout.total, dtype=torch.float32, device=dev)
        scratch = torch.zeros(1, dtype=torch.float64, device=dev)
        # grad_scale is a device scalar in general; fold it in with one multiply after the sweep
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().clasr_cl_penalty_value_grad(
                fp.theta.data_ptr(), ctx.star.data_ptr(), ctx.omega.data_ptr(), fp.layout.total, 1.0,
                scratch.data_ptr(), g.data_ptr(), _lib.stream_ptr(dev)), "cl_penalty_value_grad")
        g.mul_(grad_out.to(torch.float32))
        views = fp.layout.views(g)
        return (None, None, None) + tuple(views[n] for n in fp.layout.names)


def penalty(model, main_importance, prev_params) -> torch.Tensor:
    """Reference signature (cl_baseline_mas.py:70-75): scalar tensor, differentiable w.r.t. the parameters."""
    fp = flat_params(model)
    omega = as_flat(main_importance, fp.layout)
    star = as_flat(prev_params, fp.layout)
    return _MasPenaltyFn.apply(fp, omega.flat, star.flat, *fp.params)


def penalty_into_grads(model, main_importance, prev_params, mas_lambda: float, grad_scale: float = 1.0) -> torch.Tensor:
    """Fast path: returns the penalty value (device fp64 scalar) and ADDS ``grad_scale * mas_lambda *
    d(penalty)/d(theta)`` straight into the model's flat gradient buffer in the same sweep.

    The reference gets this gradient through autograd (cl_baseline_mas.py:231-240: ``loss += mas_lambda * penalty``
    then ``backward()``).  Writing into ``.grad`` directly bypasses autograd, so anything that multiplies the loss
    before ``backward()`` must be passed as ``grad_scale``: under ``config.mixed_precision`` the reference calls
    ``scaler.scale(loss).backward()``, i.e. every gradient carries the GradScaler's factor until ``unscale_`` —
    pass ``grad_scale=scaler.get_scale()`` (a host float; that read is the scaler's own sync) or use ``penalty()``,
    the autograd path, which needs no such care."""
    fp = flat_params(model)
    omega = as_flat(main_importance, fp.layout)
    star = as_flat(prev_params, fp.layout)
    if not fp.grads_are_flat():
        fp.gather_grads()
        fp.bind_grads(zero=False)
    value = torch.zeros(1, dtype=torch.float64, device=fp.device)
    fp.ensure_theta_views()
    with torch.cuda.device(fp.device):
        _lib.check(_lib.lib().clasr_cl_penalty_value_grad(
            fp.theta.data_ptr(), star.flat.data_ptr(), omega.flat.data_ptr(), fp.layout.total,
            float(mas_lambda) * float(grad_scale), value.data_ptr(), fp.grad.data_ptr(), _lib.stream_ptr(fp.device)), "cl_penalty_value_grad")
    return value


def mas_accumulate(importance: FlatDict, model) -> None:
    """importance[n] += |p.grad| for trainable params with a gradient (cl_baseline_mas.py:267-270)."""
    if not isinstance(importance, FlatDict) or not importance.is_intact():
        raise TypeError("mas_accumulate: `importance` must come from get_zero_params()")
    fp = flat_params(model)
    g = fp.gather_grads()
    with torch.cuda.device(fp.device):
        _lib.check(_lib.lib().clasr_cl_mas_accum(importance.flat.data_ptr(), g.data_ptr(), fp.layout.total,
                                                 _lib.stream_ptr(fp.device)), "cl_mas_accum")


def mas_finalise(importance: FlatDict, n_batches) -> FlatDict:
    """importance /= len(dataloader); main_importance = importance (overwrite) (cl_baseline_mas.py:283-287)."""
    dev = importance.flat.device
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().clasr_cl_scale_merge(importance.flat.data_ptr(), importance.flat.data_ptr(),
                                                   importance.layout.total, float(n_batches), 0.0, 1,
                                                   _lib.stream_ptr(dev)), "cl_scale_merge")
    return importance
