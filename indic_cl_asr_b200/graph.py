"""CUDA-graph capture of a whole training step (forward + backward + regulariser sweep + gradient all-reduce).

The reference launches every kernel of its step from Python and synchronises the host several times per step
(hybrid_rnnt_ctc_models.py:862-924: six ``empty_cache()`` calls and four ``.item()``; gpu_rnnt.py:229: a stream
synchronise per sub-batch).  The step built from this package never needs the host — all sizes that depend on the data
(tile tables, valid-cell counts, scales) stay on the device — so it can be recorded ONCE into a CUDA graph and replayed:
about 80 kernel launches, two streams (the CTC branch forks off the joint branch) and the NCCL all-reduce become one
``cudaGraphLaunch``.  That matters when the per-GPU batch is small (batch-sharded data parallelism at B_local = 4:
~1.5 ms of GPU work per step against ~2 ms of Python / launch overhead).

Constraints (the usual ones of whole-network capture): tensor shapes and the addresses of the step's inputs are fixed —
feed new data by copying into the tensors passed as ``static_inputs``; host-side randomness (the joint's dropout seed)
is frozen at capture time.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch

__all__ = ["GraphedStep"]


class GraphedStep:
    """``GraphedStep(fn, static_inputs)`` records ``fn(*static_inputs)`` (any Python callable that enqueues CUDA work:
    typically forward, ``backward()``, regulariser sweep, all-reduce) after ``warmup`` eager runs on the capture stream;
    ``replay()`` launches the recorded graph on the CURRENT stream and returns ``fn``'s outputs (static tensors, valid
    until the next replay).  ``pool``: share another GraphedStep's memory pool (``other.pool()``) when several graphs
    never run concurrently (e.g. the two input buffers of a prefetching loop)."""

    def __init__(self, fn: Callable, static_inputs: Sequence[torch.Tensor], warmup: int = 3, pool=None,
                 device: Optional[torch.device] = None):
        self.fn = fn
        self.static_inputs = tuple(static_inputs)
        dev = device if device is not None else self.static_inputs[0].device
        if dev.type != "cuda":
            raise RuntimeError("GraphedStep: CUDA tensors only")
        self.device = dev
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            # allocator, lazily built device tables, autograd nodes: all in steady state before the capture
            # (warmup = 0: the caller has already run fn eagerly, e.g. a second graph over another input set)
            for _ in range(max(0, warmup)):
                fn(*self.static_inputs)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        # the warm-up's transient buffers (tens of GB of GEMM scratch at the large configs) sit in the caching allocator;
        # the capture allocates its own copies from a private pool, so hand the cached ones back first
        torch.cuda.empty_cache()
        kw = {} if pool is None else {"pool": pool}
        with torch.cuda.graph(self.graph, stream=side, **kw):
            self.outputs = fn(*self.static_inputs)
        torch.cuda.synchronize(dev)

    def pool(self):
        return self.graph.pool()

    def replay(self):
        self.graph.replay()
        return self.outputs

    def __call__(self, *inputs):
        """Copy ``inputs`` into the static buffers (device-to-device or pinned host-to-device, asynchronous) and replay."""
        if len(inputs) != len(self.static_inputs):
            raise ValueError("GraphedStep: wrong number of inputs")
        with torch.no_grad():
            for dst, src in zip(self.static_inputs, inputs):
                if src is not dst:
                    dst.copy_(src, non_blocking=True)
        return self.replay()
