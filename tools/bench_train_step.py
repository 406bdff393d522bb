"""Times the walk-training step (BASELINE.json configs[2]: StyleGAN2-1024 fwd+bwd + ResNet-50 regressor, BCE loss,
batch 16/GPU) and its parts on one GPU, or data-parallel under torchrun.  Prints one JSON line on rank 0."""
import argparse
import json
import os
os.environ.setdefault("L2I_ALLOW_RANDOM_INIT", "1")   # synthetic-weight benchmark
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--no-regressor", action="store_true", help="replace ResNet-50 by a tiny conv regressor (isolates G)")
    a = ap.parse_args()
    from latent2im_b200 import parallel
    from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
    from latent2im_b200.graphs.stylegan_v2_real.transform_base import WalkLinearMultiW
    from latent2im_b200.synthetic import load_synthetic, synthetic_walk_w, synthetic_z
    from latent2im_b200.train_step import WalkTrainer

    rank, world, local = parallel.world_info()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    gen = load_synthetic(Generator(a.size, 512, 8), seed=0).to(dev).eval()
    gen.set_native(dtype=torch.bfloat16 if a.dtype == "bf16" else torch.float32, max_batch=a.batch)
    if a.no_regressor:
        reg = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, stride=4, padding=1), torch.nn.Tanh(), torch.nn.AdaptiveAvgPool2d(1),
                                  torch.nn.Flatten(), torch.nn.Linear(8, 40), torch.nn.Sigmoid())
    else:
        import torchvision
        torch.manual_seed(1)
        reg = torchvision.models.resnet50(weights=None)
        reg.fc = torch.nn.Linear(2048, 40)
        reg = torch.nn.Sequential(reg, torch.nn.Sigmoid())
    reg = reg.to(dev).eval().to(memory_format=torch.channels_last)
    np.random.seed(0)
    walk = WalkLinearMultiW(512, gen.log_size - 2, 1, ["Smiling"]).to(dev)
    with torch.no_grad():
        walk.w.copy_(synthetic_walk_w(1, gen.n_latent, 512, seed=0).to(dev))
    trainer = WalkTrainer(gen, walk, reg, [31], lr=1e-4)
    zg = synthetic_z(world * a.batch, seed=0)
    z = torch.tensor(zg[parallel.shard_rows(world * a.batch, rank, world)], dtype=torch.float32, device=dev)
    target = torch.full((a.batch, 1), float(np.random.RandomState(0).uniform(0, 1)), device=dev)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=not a.no_regressor):
        ms_step = timed(lambda: trainer.step(z, target), a.steps, a.warmup)
    # parts: G inference forward, G training forward + backward (no regressor)
    n = gen.n_latent
    with torch.no_grad():
        w = gen.style(z)
        lat = w[:, None, :].repeat(1, n, 1)
        ms_fwd = timed(lambda: gen(lat, input_is_latent=True), a.steps, a.warmup)
    probe = torch.randn(a.batch, 3, a.size, a.size, device=dev)

    def fwd_bwd():
        l = lat.clone().requires_grad_(True)
        img, _ = gen(l, input_is_latent=True)
        img.backward(probe)

    ms_fb = timed(fwd_bwd, a.steps, a.warmup)
    if rank == 0:
        print(json.dumps({"workload": f"train.py walk step, StyleGAN2-{a.size}, batch {a.batch}/GPU, {'tiny regressor' if a.no_regressor else 'ResNet-50 (bf16 autocast, channels_last)'}",
                          "n_gpus": world, "ms_per_step": ms_step, "samples_per_s": world * a.batch / ms_step * 1e3,
                          "g_forward_inference_ms": ms_fwd, "g_forward_training_plus_backward_ms": ms_fb,
                          "allreduce_bytes": trainer.last_allreduce_bytes, "dtype": a.dtype}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
