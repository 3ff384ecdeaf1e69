"""Parameter / gradient dict helpers with the reference's names and semantics (utils.py:246-321),
backed by flat buffers (cl/flat.py)."""
from __future__ import annotations

import torch

from .flat import FlatDict, flat_params

__all__ = ["freeze_layer", "get_params", "get_params_clone", "get_zero_params", "get_grads", "set_grads"]


def freeze_layer(model, num_layers):
    """utils.py:246-263: everything frozen except encoder.layers[i > num_layers], decoder, ctc_decoder, joint."""
    for param in model.parameters():
        param.requires_grad = False
    for i, layer in enumerate(model.encoder.layers):
        if i > num_layers:
            for param in layer.parameters():
                param.requires_grad = True
    for param in model.decoder.parameters():
        param.requires_grad = True
    for param in model.ctc_decoder.parameters():
        param.requires_grad = True
    for param in model.joint.parameters():
        param.requires_grad = True


def get_params(model) -> FlatDict:
    """name -> param.data for trainable params (utils.py:273-282); values alias the live parameters."""
    return flat_params(model).theta_dict()


def get_params_clone(model) -> FlatDict:
    """utils.py:284-293 — one snapshot kernel over the flat buffer instead of one clone per tensor."""
    return flat_params(model).snapshot()


def get_zero_params(model, device=None) -> FlatDict:
    """utils.py:295-302."""
    fd = flat_params(model).zeros()
    if device is not None and torch.device(device) != fd.flat.device:
        raise ValueError("get_zero_params: regulariser state lives on the model's device")
    return fd


def get_grads(model):
    """utils.py:305-313: every parameter whose .grad is not None (frozen/unused ones are skipped)."""
    out = {}
    for name, param in model.named_parameters():
        if param.grad is not None:
            out[name] = param.grad
    return out


def set_grads(model, grad_dict):
    """utils.py:316-321: param.grad = grad_dict[name], else None — called BEFORE loss.backward() so autograd
    accumulates on top of the penalty gradient (cl_baseline_ewc.py:228-240)."""
    for name, param in model.named_parameters():
        if name in grad_dict:
            param.grad = grad_dict[name]
        else:
            param.grad = None
