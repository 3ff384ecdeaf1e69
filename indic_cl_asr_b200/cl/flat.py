"""Flat fp32 parameter / gradient / importance buffers behind the reference's dict-of-tensors hooks.

The reference keeps EWC/MAS state as ``{name: tensor}`` dicts over ``named_parameters()`` filtered by
``requires_grad`` (utils.py:273-321) and walks them in Python.  Here every such dict is a ``FlatDict``:
ordinary dict semantics, but all values are views into ONE contiguous fp32 buffer (each tensor padded to
4 floats), so that each hook is a single streaming kernel over HBM and the multi-GPU exchange is a single
all-reduce.  Buffers live in HBM for the whole run: theta, grad, theta*, F/Omega = 4 x 4 B/param
(2.1 GB for the 129 M-parameter model — trivial next to 180 GB).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np
import torch

from .. import _lib

SWEEP_CHUNK = 8192  # CLASR_SWEEP_CHUNK
_ITEM_DTYPE = np.dtype([("start", "<i8"), ("len", "<i4"), ("seg", "<i4")])


class Layout:
    """name -> (offset, numel, shape) in named_parameters() order; offsets are multiples of 4 floats."""

    def __init__(self, named_shapes: Iterable[Tuple[str, torch.Size]]):
        self.names: List[str] = []
        self.offsets: List[int] = []
        self.numels: List[int] = []
        self.shapes: List[torch.Size] = []
        off = 0
        for name, shape in named_shapes:
            n = int(np.prod(shape)) if len(shape) else 1
            self.names.append(name)
            self.offsets.append(off)
            self.numels.append(n)
            self.shapes.append(torch.Size(shape))
            off += (n + 3) // 4 * 4
        self.total = off
        self._dev_cache: Dict[str, Tuple[torch.Tensor, torch.Tensor, int]] = {}

    def key(self):
        return (tuple(self.names), tuple(self.numels))

    def sweep_items(self) -> np.ndarray:
        items = []
        for seg, (off, n) in enumerate(zip(self.offsets, self.numels)):
            for s in range(0, n, SWEEP_CHUNK):
                items.append((off + s, min(SWEEP_CHUNK, n - s), seg))
        return np.array(items, dtype=_ITEM_DTYPE)

    def device_tables(self, device) -> Tuple[torch.Tensor, torch.Tensor, int]:
        """(items [n_items] as raw bytes on device, seg_numel int64 [n_seg], n_items)."""
        k = str(device)
        if k not in self._dev_cache:
            items = self.sweep_items()
            raw = torch.from_numpy(items.view(np.uint8).copy()).to(device)
            numel = torch.tensor(self.numels, dtype=torch.int64, device=device)
            self._dev_cache[k] = (raw, numel, len(items))
        return self._dev_cache[k]

    def views(self, flat: torch.Tensor) -> "OrderedDict[str, torch.Tensor]":
        out = OrderedDict()
        for name, off, n, shape in zip(self.names, self.offsets, self.numels, self.shapes):
            out[name] = flat[off:off + n].view(shape)
        return out


class FlatDict(OrderedDict):
    """dict name -> view, plus ``.flat`` (the backing buffer) and ``.layout``."""

    def __init__(self, layout: Layout, flat: torch.Tensor):
        super().__init__(layout.views(flat))
        self.layout = layout
        self.flat = flat

    def is_intact(self) -> bool:
        """True while every value still aliases its slot of the flat buffer (a user may rebind keys)."""
        base = self.flat.data_ptr()
        if list(self.keys()) != self.layout.names:
            return False
        for (name, t), off in zip(self.items(), self.layout.offsets):
            if t is None or t.data_ptr() != base + 4 * off or not t.is_contiguous():
                return False
        return True


def _alloc(layout: Layout, device, zero=True) -> torch.Tensor:
    return (torch.zeros if zero else torch.empty)(layout.total, dtype=torch.float32, device=device)


def as_flat(d: Dict[str, torch.Tensor], layout: Optional[Layout] = None) -> FlatDict:
    """Adopt a plain ``{name: tensor}`` dict: returns it unchanged when already flat and intact, else packs
    a copy into a new flat buffer (one torch.cat; the arithmetic still runs in the sweep kernels)."""
    if isinstance(d, FlatDict) and d.is_intact() and (layout is None or d.layout.key() == layout.key()):
        return d
    if layout is None:
        layout = Layout((k, v.shape) for k, v in d.items())
    first = next(iter(d.values()))
    _lib.require_cuda(first, "regulariser state")
    flat = _alloc(layout, first.device)
    fd = FlatDict(layout, flat)
    for name in layout.names:
        fd[name].copy_(d[name])
    return fd


class FlatParams:
    """Re-homes a module's trainable parameters (and their .grad) into flat buffers."""

    def __init__(self, model: torch.nn.Module):
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        if not named:
            raise ValueError("model has no trainable parameters")
        dev = named[0][1].device
        _lib.require_cuda(named[0][1], "model parameters")
        for n, p in named:
            if p.dtype != torch.float32:
                raise TypeError(f"parameter {n} is {p.dtype}; the regulariser sweeps are fp32")
            if p.device != dev:
                raise ValueError("all trainable parameters must live on one device")
        self.model = model
        self.layout = Layout((n, p.shape) for n, p in named)
        self.params = [p for _, p in named]
        self.theta = _alloc(self.layout, dev)
        self.grad = _alloc(self.layout, dev)
        tv = self.layout.views(self.theta)
        with torch.no_grad():
            for (n, p) in named:
                tv[n].copy_(p.data)
                p.data = tv[n]
        self._grad_views = self.layout.views(self.grad)
        self.device = dev

    # -- views ---------------------------------------------------------------------------------
    def theta_dict(self) -> FlatDict:
        self.ensure_theta_views()
        return FlatDict(self.layout, self.theta)

    def grad_dict(self) -> FlatDict:
        return FlatDict(self.layout, self.grad)

    def zeros(self) -> FlatDict:
        return FlatDict(self.layout, _alloc(self.layout, self.device))

    def snapshot(self) -> FlatDict:
        """theta* <- theta (get_params_clone, utils.py:284-293) in one kernel."""
        self.ensure_theta_views()
        out = _alloc(self.layout, self.device, zero=False)
        with torch.cuda.device(self.device):
            st = _lib.lib().clasr_cl_snapshot(out.data_ptr(), self.theta.data_ptr(), self.layout.total,
                                              _lib.stream_ptr(self.device))
        _lib.check(st, "cl_snapshot")
        return FlatDict(self.layout, out)

    def ensure_theta_views(self) -> None:
        """Re-adopt parameters whose .data was rebound (e.g. load_state_dict keeps views; .to() does not)."""
        base = self.theta.data_ptr()
        with torch.no_grad():
            for p, off, (name, view) in zip(self.params, self.layout.offsets, self.layout.views(self.theta).items()):
                if p.data.data_ptr() != base + 4 * off:
                    view.copy_(p.data)
                    p.data = view

    def bind_grads(self, zero: bool = True) -> None:
        """Point every trainable parameter's .grad at its slot of the flat gradient buffer (use instead of
        optimizer.zero_grad(): autograd then accumulates straight into the flat buffer)."""
        if zero:
            self.grad.zero_()
        for p, (name, view) in zip(self.params, self._grad_views.items()):
            p.grad = view

    def grads_are_flat(self) -> bool:
        base = self.grad.data_ptr()
        for p, off in zip(self.params, self.layout.offsets):
            if p.grad is None or p.grad.data_ptr() != base + 4 * off:
                return False
        return True

    def gather_grads(self) -> torch.Tensor:
        """Flat gradient buffer reflecting the parameters' current .grad (None -> zeros, as a parameter
        without gradient contributes nothing to Fisher / Omega: utils.py:305-313)."""
        if self.grads_are_flat():
            return self.grad
        with torch.no_grad():
            for p, (name, view) in zip(self.params, self._grad_views.items()):
                if p.grad is None:
                    view.zero_()
                elif p.grad.data_ptr() != view.data_ptr():
                    view.copy_(p.grad)
        return self.grad


def flat_params(model: torch.nn.Module) -> FlatParams:
    """The FlatParams attached to ``model`` (created on first use)."""
    fp = getattr(model, "_clasr_flat", None)
    if fp is None or [id(p) for p in fp.params] != [id(p) for _, p in model.named_parameters() if p.requires_grad]:
        fp = FlatParams(model)
        object.__setattr__(model, "_clasr_flat", fp)
    return fp
