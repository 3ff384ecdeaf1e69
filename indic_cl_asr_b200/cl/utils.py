"""Parameter / gradient dict helpers with the reference's names and semantics (utils.py:246-321),
backed by flat buffers (cl/flat.py)."""
from __future__ import annotations

import torch

from .flat import FlatDict, flat_params

__all__ = ["freeze_layer", "get_params", "get_params_clone", "get_zero_params", "get_grads", "set_grads",
           "save_cl_state", "load_cl_state"]


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
    """utils.py:295-302: zeros_like(param.data).to(device) per trainable tensor.  The flat buffer lives on the model's
    device (where the sweeps run); asking for another device gets a flat copy there, as the reference's ``.to(device)``
    would, and the hooks re-home it (``as_flat``) when it meets the model again."""
    fd = flat_params(model).zeros()
    if device is not None:
        dev = torch.device(device)
        if dev.type == "cuda" and dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else dev
        if dev != fd.flat.device:
            return FlatDict(fd.layout, fd.flat.to(dev))
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


# ------------------------------------------------------------------------------------------------
# Continual-learning state on disk.  The reference keeps theta*, Fisher / Omega only in process memory
# (cl_baseline_ewc.py:267-282, cl_baseline_mas.py:283-288; utils.save_model :265-271 stores weights only), so a
# crashed 12-language run restarts from language 0.  With flat buffers the whole state is a handful of 1-D tensors.
# ------------------------------------------------------------------------------------------------
_CL_STATE_VERSION = 1


def save_cl_state(path, checkpoint=None, importance=None, **extra) -> None:
    """Write ``{theta* (checkpoint), F or Omega (importance), names/shapes, extra scalars}`` to ``path``.
    ``checkpoint`` / ``importance`` are the dicts returned by get_params_clone / get_zero_params (or plain dicts)."""
    from .flat import as_flat

    ref = checkpoint if checkpoint is not None else importance
    if ref is None:
        raise ValueError("save_cl_state: nothing to save")
    lay = as_flat(ref).layout
    blob = {"version": _CL_STATE_VERSION, "names": list(lay.names), "shapes": [tuple(s) for s in lay.shapes],
            "extra": dict(extra)}
    for key, d in (("checkpoint", checkpoint), ("importance", importance)):
        blob[key] = None if d is None else as_flat(d, lay).flat.detach().cpu()
    torch.save(blob, path)


def load_cl_state(path, model):
    """Inverse of save_cl_state for ``model`` (same trainable parameters in the same order, utils.py:273-282).
    Returns ``(checkpoint, importance, extra)`` as FlatDicts on the model's device (None where not saved)."""
    fp = flat_params(model)
    blob = torch.load(path, map_location="cpu", weights_only=False)
    if blob.get("version") != _CL_STATE_VERSION:
        raise ValueError(f"load_cl_state: unsupported version {blob.get('version')}")
    if list(blob["names"]) != list(fp.layout.names) or [tuple(s) for s in blob["shapes"]] != [tuple(s) for s in fp.layout.shapes]:
        raise ValueError("load_cl_state: the saved parameter layout does not match the model's trainable parameters")
    out = []
    for key in ("checkpoint", "importance"):
        t = blob[key]
        if t is None:
            out.append(None)
        else:
            if t.numel() != fp.layout.total:
                raise ValueError(f"load_cl_state: `{key}` has {t.numel()} elements, expected {fp.layout.total}")
            out.append(FlatDict(fp.layout, t.to(fp.device, dtype=torch.float32).contiguous()))
    return out[0], out[1], blob["extra"]
