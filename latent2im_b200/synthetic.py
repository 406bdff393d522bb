"""Deterministic synthetic inputs: generator weights, noise, latents, walks.

There is no network for checkpoints, so benchmarks and parity tests run on random-init weights of
the real architecture.  ``synthetic_state_dict`` fills a rosinality-format ``state_dict`` *by key
name, in sorted key order, from one seeded CPU generator*, so the reference's own ``Generator``
(on the GPU box) and this package's ``Generator`` receive bit-identical tensors regardless of
module construction order.  Unlike plain random init it makes the noise weights and every bias
non-zero - with the stock init those paths are numerically invisible (networks.py:279,
op/fused_act.py:77).
"""
from __future__ import annotations

import numpy as np
import torch


def synthetic_state_dict(shapes: dict, seed: int = 0, lr_mlp: float = 0.01, rgb_gain: float = 1.0) -> dict:
    """``shapes``: key -> shape (from any Generator's ``state_dict()``).  FIR kernels
    (``*.kernel``) are left out - they are constants of the architecture.

    ``rgb_gain`` scales the ToRGB conv weights and biases.  With the unit-variance recipe (1.0, the one the golden
    fixtures were generated with) a random-init generator emits images spanning about +-7, seven times the [-1, 1]
    range of a trained one; 0.25 brings the synthetic images into that range, which is what an absolute
    (peak-to-peak 2) PSNR gate presumes."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    out = {}
    for key in sorted(shapes):
        shape = tuple(shapes[key])
        if key.endswith(".kernel"):
            continue
        r = torch.randn(shape, generator=g, dtype=torch.float32)
        if key.startswith("style.") and key.endswith(".weight"):
            v = r / lr_mlp                      # EqualLinear(..., lr_mul=lr_mlp): randn / lr_mul
        elif key.startswith("style.") and key.endswith(".bias"):
            v = 0.1 * r / lr_mlp                # effective bias = bias * lr_mul ~ 0.1 * N(0,1)
        elif key.endswith(".modulation.bias"):
            v = 1.0 + 0.1 * r                   # bias_init = 1
        elif key.endswith(".noise.weight"):
            v = 0.1 * r
        elif key.endswith(".bias") and "to_rgb" in key:
            v = 0.1 * r * rgb_gain
        elif key.endswith(".activate.bias"):
            v = 0.1 * r
        elif "to_rgb" in key and key.endswith(".conv.weight"):
            v = r * rgb_gain
        else:                                   # conv / modulation / const input / noise buffers
            v = r
        out[key] = v
    return out


def synthetic_discriminator_state_dict(shapes: dict, seed: int = 0) -> dict:
    """Same idea for a ``Discriminator.state_dict()``: unit-normal weights, 0.1 * N(0, 1) biases, FIR kernels left out."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    out = {}
    for key in sorted(shapes):
        if key.endswith(".kernel"):
            continue
        r = torch.randn(tuple(shapes[key]), generator=g, dtype=torch.float32)
        out[key] = 0.1 * r if key.endswith(".bias") else r
    return out


def load_synthetic(generator, seed: int = 0, rgb_gain: float = 1.0):
    """Overwrites ``generator``'s parameters and noise buffers in place with the synthetic recipe."""
    sd = generator.state_dict()
    syn = synthetic_state_dict({k: v.shape for k, v in sd.items()}, seed, getattr(generator, "lr_mlp", 0.01), rgb_gain)
    with torch.no_grad():
        for k, v in syn.items():
            sd[k].copy_(v.to(sd[k].device))
    return generator


def synthetic_noise(num_layers: int, batch: int, seed: int = 2, device="cpu"):
    """Explicit per-layer noise list ([batch, 1, H, W] float32) for parity runs."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    out = []
    for i in range(num_layers):
        r = 2 ** ((i + 5) // 2)
        out.append(torch.randn(batch, 1, r, r, generator=g, dtype=torch.float32).to(device))
    return out


def synthetic_z(batch: int, seed: int = 0, dim_z: int = 512) -> np.ndarray:
    """Same stream as the reference's graph_util.z_sample (graph_util.py:5-8)."""
    return np.random.RandomState(seed).randn(batch, dim_z)


def synthetic_walk_w(n_attr: int, n_latent: int, dim: int = 512, seed: int = 0) -> torch.Tensor:
    """WalkLinearMultiW init (transform_base.py:147) from a *seeded* numpy stream."""
    rs = np.random.RandomState(seed)
    return torch.tensor(rs.normal(0.0, 0.02, [n_attr, n_latent, dim]), dtype=torch.float32)
