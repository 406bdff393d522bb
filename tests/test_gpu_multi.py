"""2 GPUs, NCCL: the data-parallel walk-training step (latents sharded by rows, ONE all-reduce of the flat walk gradient)
leaves the same walk parameters as a single process stepping on the global batch.  Skipped on a 1-GPU box; run with
``gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu``."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _build(device, batch):
    from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
    from latent2im_b200.graphs.stylegan_v2_real.transform_base import WalkLinearMultiW
    from latent2im_b200.synthetic import load_synthetic, synthetic_walk_w
    from latent2im_b200.train_step import WalkTrainer
    size, dim = 16, 32
    gen = load_synthetic(Generator(size, dim, 2), seed=0).to(device).eval()
    gen.set_native(dtype=torch.float32, max_batch=batch)
    torch.manual_seed(1)
    reg = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.Tanh(), torch.nn.AdaptiveAvgPool2d(1), torch.nn.Flatten(),
                              torch.nn.Linear(4, 5), torch.nn.Sigmoid()).to(device)
    import numpy as np
    np.random.seed(0)
    walk = WalkLinearMultiW(dim, gen.log_size - 2, 1, ["Smiling"]).to(device)
    with torch.no_grad():
        walk.w.copy_(synthetic_walk_w(1, gen.n_latent, dim, seed=0).to(device))
    return gen, walk, WalkTrainer(gen, walk, reg, [2], lr=1e-2)


def _run(trainer, gen, z, target, noise, steps):
    orig = gen.forward
    gen.forward = lambda styles, **kw: orig(styles, noise=noise, **kw)
    for _ in range(steps):
        trainer.step(z, target)
    gen.forward = orig


def _worker(rank, world, port, steps, out):
    import torch.distributed as dist
    from latent2im_b200 import parallel
    from latent2im_b200.synthetic import synthetic_noise, synthetic_z
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        b, gb = 2, 2 * world
        gen, walk, trainer = _build(dev, b)
        rows = parallel.shard_rows(gb, rank, world)
        z = torch.tensor(synthetic_z(gb, 0, 32), dtype=torch.float32)[rows].to(dev)
        target = torch.full((b, 1), 0.8, device=dev)
        noise = [n[rows].to(dev) for n in synthetic_noise(gen.num_layers, gb)]
        _run(trainer, gen, z, target, noise, steps)
        if rank == 0:
            out["w"] = walk.w.detach().cpu()
            out["bytes"] = trainer.last_allreduce_bytes
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_gpu_nccl_step_equals_single_gpu_global_batch():
    import torch.multiprocessing as mp
    from latent2im_b200.synthetic import synthetic_noise, synthetic_z
    steps, world = 2, 2
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), steps, out), nprocs=world, join=True)
    dev = torch.device("cuda", 0)
    gb = 2 * world
    gen, walk, trainer = _build(dev, gb)
    z = torch.tensor(synthetic_z(gb, 0, 32), dtype=torch.float32).to(dev)
    noise = [n.to(dev) for n in synthetic_noise(gen.num_layers, gb)]
    _run(trainer, gen, z, torch.full((gb, 1), 0.8, device=dev), noise, steps)
    assert out["bytes"] == walk.w.numel() * 4
    ref = walk.w.detach().cpu()
    assert (out["w"] - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item())
