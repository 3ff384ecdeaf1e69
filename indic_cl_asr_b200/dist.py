"""Batch-sharded data parallelism for the hot path (one process per GPU, torch.distributed/NCCL).

The reference wraps the model in DDP but calls ``model.module.training_step`` (cl_baseline_ewc.py:225), so no
gradient is ever exchanged; N-GPU behaviour is therefore defined here as "equals the 1-GPU
result on the concatenated batch".  Utterances are independent units: each rank takes a contiguous slice
of the batch, weights and regulariser state are replicated, and the only data-path exchange is ONE sum
all-reduce per step over the flat fp32 gradient buffer (cl/flat.py) plus, once per task, one over the
flat Fisher / Omega buffer and the sample count.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "shard_batch", "local_loss_scale", "allreduce_flat_", "allreduce_importance_"]


def _ws(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_bounds(batch_size: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [begin, end) of the batch for `rank` (first B % W ranks get one extra)."""
    base, extra = divmod(batch_size, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_batch(tensors, rank: Optional[int] = None, world_size: Optional[int] = None, group=None):
    r, w = _ws(group)
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    b, e = shard_bounds(int(tensors[0].shape[0]), rank, world_size)
    return tuple(t[b:e] for t in tensors)


def local_loss_scale(local_batch: int, global_batch: int) -> float:
    """Weight of the local mean_batch loss so that the SUM all-reduce of gradients reproduces the gradient of
    the global mean over the concatenated batch, even when shards are ragged."""
    return float(local_batch) / float(global_batch)


def allreduce_flat_(flat: torch.Tensor, group=None, async_op: bool = False):
    """In-place sum all-reduce of a flat buffer (gradients).  One collective per step."""
    if _ws(group)[1] == 1:
        return None
    return dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def allreduce_importance_(flat: torch.Tensor, count: torch.Tensor, group=None) -> None:
    """Per task: sum the flat Fisher / Omega accumulator and the sample (EWC) or batch (MAS) count across ranks
    BEFORE the `/= count` and gamma-merge (cl_baseline_ewc.py:267-280, cl_baseline_mas.py:283-287)."""
    if _ws(group)[1] == 1:
        return
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(count, op=dist.ReduceOp.SUM, group=group)
