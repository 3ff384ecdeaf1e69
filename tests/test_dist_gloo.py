"""world_size-2 gloo test of the host-side multi-GPU logic (indic_cl_asr_b200/dist.py): batch sharding +
loss scaling + one flat sum all-reduce must reproduce the single-process gradient on the concatenated batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from indic_cl_asr_b200.dist import allreduce_flat_, allreduce_importance_, local_loss_scale, shard_bounds


def test_shard_bounds_cover_batch():
    for B in (1, 7, 32, 33):
        for W in (1, 2, 4, 8):
            spans = [shard_bounds(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    w = torch.randn(6, 3, requires_grad=True)
    x = torch.randn(B, 6)
    y = torch.randn(B, 3)
    b, e = shard_bounds(B, rank, world)
    # local mean_batch loss, weighted so that the SUM all-reduce equals the global mean's gradient
    loss = ((x[b:e] @ w - y[b:e]) ** 2).sum(1).mean() * local_loss_scale(e - b, B)
    loss.backward()
    flat = w.grad.detach().clone().flatten()
    allreduce_flat_(flat)
    # per-task importance exchange: sum of accumulators and of counts
    fisher = (w.grad.detach() ** 2).flatten().clone()
    count = torch.tensor([float(e - b)])
    allreduce_importance_(fisher, count)
    if rank == 0:
        out["grad"] = flat.numpy()
        out["count"] = float(count)
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_matches_single_process():
    B = 7  # ragged shards: 4 + 3
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), B, out), nprocs=2, join=True)
    torch.manual_seed(0)
    w = torch.randn(6, 3, requires_grad=True)
    x = torch.randn(B, 6)
    y = torch.randn(B, 3)
    ((x @ w - y) ** 2).sum(1).mean().backward()
    assert np.allclose(out["grad"], w.grad.flatten().numpy(), rtol=1e-5, atol=1e-6)
    assert out["count"] == B


def _worker_strong(rank, world, port, B, out):
    """bench.py's strong-scaling step on 2 ranks: every rank pre-loads 1/N of the EWC penalty gradient, back-propagates
    its shard's mean loss weighted by B_local/B, then ONE sum all-reduce of the flat buffer."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    w = torch.randn(6, 3, requires_grad=True)
    x, y = torch.randn(B, 6), torch.randn(B, 3)
    star, fish = torch.randn(6, 3), torch.rand(6, 3)
    b, e = shard_bounds(B, rank, world)
    w.grad = (10.0 / world) * 2 * fish * (w.detach() - star)          # get_penalty_grads with e_lambda / N, set_grads
    loss = ((x[b:e] @ w - y[b:e]) ** 2).sum(1).mean() * local_loss_scale(e - b, B)
    loss.backward()                                                   # accumulates on top of the pre-loaded penalty
    flat = w.grad.detach().clone().flatten()
    allreduce_flat_(flat)
    if rank == 0:
        out["grad"] = flat.numpy()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_strong_scaling_step_matches_single_process_ewc_step():
    B = 7
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_strong, args=(2, _free_port(), B, out), nprocs=2, join=True)
    torch.manual_seed(0)
    w = torch.randn(6, 3, requires_grad=True)
    x, y = torch.randn(B, 6), torch.randn(B, 3)
    star, fish = torch.randn(6, 3), torch.rand(6, 3)
    w.grad = 10.0 * 2 * fish * (w.detach() - star)                    # cl_baseline_ewc.py:74, 228-231
    ((x @ w - y) ** 2).sum(1).mean().backward()                       # :240
    assert np.allclose(out["grad"], w.grad.flatten().numpy(), rtol=1e-5, atol=1e-6)
