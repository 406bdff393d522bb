"""Data-parallel plumbing of the walk-training step: one process per GPU, batch-of-latents sharding,
and the single collective of the path - the all-reduce of the (tiny) walk-parameter gradient.

The reference is single-GPU; every latent sample is independent through mapping,
walk, G and the regressor (section 8e), so the global batch ``world * b`` is split by rows and the only
cross-rank dependency is the mean of the walk gradient.  All reference losses are batch means
(transform_base.py:412-414), so averaging the per-rank mean gradients over equal shards equals the
gradient of the global-batch mean.  Backend: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def world_info() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))


def shard_rows(n_global: int, rank: int, world: int) -> slice:
    """Rows of a global batch owned by ``rank``: contiguous equal shards, so the union over ranks is the
    single-GPU batch in order."""
    if n_global % world != 0:
        raise ValueError(f"global batch {n_global} is not divisible by world size {world}")
    b = n_global // world
    return slice(rank * b, (rank + 1) * b)


def flatten_grads(params: Iterable[torch.nn.Parameter]) -> Tuple[torch.Tensor, List[torch.nn.Parameter]]:
    """One contiguous fp32 buffer holding every parameter gradient (zeros where a gradient is missing)."""
    ps = [p for p in params if p.requires_grad]
    if not ps:
        return torch.zeros(0), ps
    dev = ps[0].device
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).to(torch.float32) for p in ps]).to(dev)
    return flat, ps


def unflatten_grads(flat: torch.Tensor, ps: List[torch.nn.Parameter]) -> None:
    off = 0
    for p in ps:
        n = p.numel()
        g = flat[off:off + n].reshape(p.shape).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n


def allreduce_mean_grads(params: Iterable[torch.nn.Parameter], group=None) -> int:
    """Averages the gradients of ``params`` over the process group with ONE all-reduce of a flat buffer
    (36 KB for the linear walk with one attribute, 8.4 MB for the MLP walk).  Returns the bytes reduced.
    A no-op outside a process group (single GPU)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    flat, ps = flatten_grads(params)
    if flat.numel() == 0:
        return 0
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(dist.get_world_size(group))
    unflatten_grads(flat, ps)
    return flat.numel() * 4


def broadcast_params(params: Iterable[torch.nn.Parameter], src: int = 0, group=None) -> None:
    """Makes every rank start from rank ``src``'s walk parameters (the reference draws them from the
    unseeded global numpy RNG, transform_base.py:147)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for p in params:
        dist.broadcast(p.data, src=src, group=group)
