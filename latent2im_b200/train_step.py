"""The walk-training step of the reference's ``train.py`` (lines 48-116 + ``optimizeParametersAll``,
transform_base.py:456-490) on the native hot path, data parallel over latents.

Per step and rank (with the dead branch removed):
    w        = G.style(z)                                  no autograd  (l2i_generator_mapping)
    alpha0   = R(G(w))[:, attr]                            no autograd  (G forward #1, inference kernels)
    eps      = target - alpha0
    w'       = walk([w] * n_latent, eps)                   autograd     (l2i_walk_* kernels)
    image    = G(w')                                       autograd     (training-mode forward, activations kept natively)
    loss     = BCE(R(image)[:, attr], target)              mean over the local batch (transform_base.py:412-424)
    loss.backward()                                        R by autograd (frozen), G by l2i_generator_backward,
                                                           walk by l2i_walk_linear_bwd / the MLP kernels
    all-reduce(mean) of the flat walk gradient             the ONLY collective (NCCL over NVLink)
    Adam(lr, betas=(0.5, 0.99)).step()                     transform_base.py:329-331
The discriminator / VGG terms (TransformGraph.optimizeParametersAll without ``--no_gan_loss --no_content_loss``) are not part of
this fused step: the discriminator runs on the repository's blur / bias-act kernels + cuDNN, VGG19 is stock PyTorch.
"""
from __future__ import annotations

import torch

from . import parallel


def bce_clamped(pred: torch.Tensor, y: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """transform_base.py:412-414 (``get_bce_loss``)."""
    return -(y * pred.clamp(min=eps).log() + (1 - y) * (1 - pred).clamp(min=eps).log()).mean()


class WalkTrainer:
    def __init__(self, generator, walk, regressor, attr_idx, lr: float = 1e-4, group=None):
        self.gen, self.walk, self.reg, self.attr_idx, self.group = generator, walk, regressor, list(attr_idx), group
        for p in self.reg.parameters():
            p.requires_grad_(False)          # frozen: data gradients only
        self.opt = torch.optim.Adam(self.walk.parameters(), lr=lr, betas=(0.5, 0.99))
        parallel.broadcast_params(self.walk.parameters(), 0, group)
        self.last_allreduce_bytes = 0

    def _preds(self, image):
        p = self.reg(image)[:, self.attr_idx]
        return p.unsqueeze(1) if p.ndim == 1 else p

    def step(self, z_local: torch.Tensor, target: torch.Tensor, layers=None) -> torch.Tensor:
        """``z_local``: this rank's rows of the global z batch [b, 512] (device); ``target``: [b, A] attribute targets.
        Returns the local loss (device scalar; no host sync here - train.py's ``.item()`` is the caller's choice)."""
        n = self.gen.n_latent
        with torch.no_grad():
            w = self.gen.style(z_local)
            img0, _ = self.gen(w[:, None, :].expand(-1, n, -1), input_is_latent=True)
            eps = target - self._preds(img0)
        self.opt.zero_grad(set_to_none=True)
        ws = self.walk([w] * n, eps, layers=layers)
        lat = torch.stack(ws, 1)
        image, _ = self.gen(lat, input_is_latent=True)
        loss = bce_clamped(self._preds(image), target)
        loss.backward()
        self.last_allreduce_bytes = parallel.allreduce_mean_grads(self.walk.parameters(), self.group)
        self.opt.step()
        return loss.detach()
