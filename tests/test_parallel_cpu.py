"""CPU, world_size 2, gloo: the N > 1 host logic of the path - latent sharding and the single collective
(flat walk-gradient all-reduce) - reproduces the single-process global-batch update exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from latent2im_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _ToyWalk(torch.nn.Module):
    """Same parameterisation as WalkLinearMultiW (transform_base.py:140-165) in plain torch, CPU."""

    def __init__(self, n_attr, n_latent, dim):
        super().__init__()
        rs = np.random.RandomState(0)
        self.w = torch.nn.Parameter(torch.tensor(rs.normal(0.0, 0.02, [n_attr, n_latent, dim]), dtype=torch.float32))
        self.b = torch.nn.Parameter(torch.zeros(dim))   # second parameter: exercises flatten / unflatten offsets

    def forward(self, w0, alpha):
        return w0[:, None, :] + torch.einsum("ba,aid->bid", alpha, self.w) + self.b


def _loss(walk, z, alpha, probe):
    # stand-in for BCE(R(G(walk(w)))): any per-sample differentiable function with a batch-mean reduction
    lat = walk(torch.tanh(z), alpha)
    return ((lat * probe).sum(dim=(1, 2)) ** 2).mean()


def _global_inputs(n, n_latent, dim, n_attr):
    z = torch.tensor(np.random.RandomState(3).randn(n, dim), dtype=torch.float32)
    alpha = torch.tensor(np.random.RandomState(4).uniform(0, 1, [n, n_attr]), dtype=torch.float32)
    probe = torch.randn(n_latent, dim, generator=torch.Generator().manual_seed(5))
    return z, alpha, probe


def _worker(rank, world, port, steps, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, n_latent, dim, n_attr = 8, 6, 16, 2
        z, alpha, probe = _global_inputs(n, n_latent, dim, n_attr)
        rows = parallel.shard_rows(n, rank, world)
        walk = _ToyWalk(n_attr, n_latent, dim)
        if rank != 0:                      # ranks start different; broadcast must fix that
            with torch.no_grad():
                walk.w.add_(1.0)
        parallel.broadcast_params(walk.parameters(), 0)
        opt = torch.optim.Adam(walk.parameters(), lr=1e-2, betas=(0.5, 0.99))
        nbytes = 0
        for _ in range(steps):
            opt.zero_grad(set_to_none=True)
            _loss(walk, z[rows], alpha[rows], probe).backward()
            nbytes = parallel.allreduce_mean_grads(walk.parameters())
            opt.step()
        if rank == 0:
            out["w"] = walk.w.detach().clone()
            out["b"] = walk.b.detach().clone()
            out["bytes"] = nbytes
    finally:
        dist.destroy_process_group()


def test_shard_rows_partition_the_global_batch():
    rows = [parallel.shard_rows(32, r, 4) for r in range(4)]
    assert [(s.start, s.stop) for s in rows] == [(0, 8), (8, 16), (16, 24), (24, 32)]
    with pytest.raises(ValueError):
        parallel.shard_rows(10, 0, 4)


def test_allreduce_is_a_noop_without_a_group():
    walk = _ToyWalk(1, 2, 4)
    walk.w.grad = torch.ones_like(walk.w)
    assert parallel.allreduce_mean_grads(walk.parameters()) == 0
    assert torch.equal(walk.w.grad, torch.ones_like(walk.w))


@pytest.mark.timeout(120)
def test_two_rank_gloo_equals_single_process_global_batch():
    steps = 3
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), steps, out), nprocs=2, join=True)
    # single process, global batch
    n, n_latent, dim, n_attr = 8, 6, 16, 2
    z, alpha, probe = _global_inputs(n, n_latent, dim, n_attr)
    walk = _ToyWalk(n_attr, n_latent, dim)
    opt = torch.optim.Adam(walk.parameters(), lr=1e-2, betas=(0.5, 0.99))
    for _ in range(steps):
        opt.zero_grad(set_to_none=True)
        _loss(walk, z, alpha, probe).backward()
        opt.step()
    assert out["bytes"] == (walk.w.numel() + walk.b.numel()) * 4      # ONE flat buffer holds every gradient
    assert torch.allclose(out["w"], walk.w.detach(), atol=1e-6, rtol=1e-5)
    assert torch.allclose(out["b"], walk.b.detach(), atol=1e-6, rtol=1e-5)
