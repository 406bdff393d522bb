"""GAN inversion into W+ on the native forward / data-gradient kernels - the second consumer of the hot path in the
reference (``BP.py::Trainer.train``, BP.py:119-171).

The reference optimises ``w`` of shape [N, n_latent, 512], initialised at ``mean_latent(4096)``, with Adam(betas=(0.5, 0.9))
through ``netG(w, input_is_latent=True)`` on ``-log_likelihood(output, batch) / (H*W*3)`` - the branch the reference
actually executes is ``-sum(diff^2, [1,2,3])`` (BP.py:84-87; its Gaussian branch is dead code) - plus a VGG16 Gram-matrix
perceptual loss whose weights need the network (``perceptual_vgg/vgg.py``).  Here the squared-error term is implemented
exactly; a perceptual term can be passed as ``extra_loss(image, target) -> [N]``.  Every step is one training-mode forward and one
``l2i_generator_backward`` - no generator weight gradient is ever formed.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch


def reconstruction_loss(output: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """``-log_likelihood(output, target) / N`` with N = H*W*3 (BP.py:84-87, 143-144): per-sample mean squared error [N]."""
    n = target.shape[2] * target.shape[3] * 3
    return (output - target).pow(2).sum(dim=[1, 2, 3]) / n


class GramPerceptualLoss(torch.nn.Module):
    """The perceptual term of BP.py:173-185: squared differences of the Gram matrices of four VGG16 activations
    (after ``features[3]``, ``[8]``, ``[15]``, ``[22]`` = relu1_2, relu2_2, relu3_3, relu4_3; perceptual_vgg/vgg.py:7-36),
    each weighted by its Gram size ``C*C``.  Stock torchvision; ``weights_path`` (a ``vgg16`` state_dict) stands in for the
    reference's ``pretrained=True`` download.  Returns one value per sample."""

    TAPS = (3, 8, 15, 22)

    def __init__(self, weights_path: str = ""):
        super().__init__()
        import torchvision
        vgg = torchvision.models.vgg16(weights=None)
        if weights_path:
            vgg.load_state_dict(torch.load(weights_path, map_location="cpu", weights_only=False))
        self.features = torch.nn.Sequential(*[torch.nn.ReLU(inplace=False) if isinstance(m, torch.nn.ReLU) else m
                                              for m in vgg.features[:self.TAPS[-1] + 1]]).eval()
        for p in self.parameters():
            p.requires_grad_(False)

    @staticmethod
    def gram(x):
        b, c, h, w = x.shape
        f = x.reshape(b, c, h * w)
        return f.bmm(f.transpose(1, 2)) / (c * h * w)

    def grams(self, x):
        out = []
        for i, m in enumerate(self.features):
            x = m(x)
            if i in self.TAPS:
                out.append(self.gram(x))
        return out

    def forward(self, image, target):
        loss = image.new_zeros(image.shape[0])
        for gi, gt in zip(self.grams(image), self.grams(target)):
            loss = loss + (gt - gi).pow(2).sum(dim=[1, 2]) * (gt.shape[1] * gt.shape[2])
        return loss


def invert(generator, target: torch.Tensor, n_loops: int = 500, lr: float = 1e-4, optim: str = "Adam",
           noise=None, mean_latent: Optional[torch.Tensor] = None, extra_loss: Optional[Callable] = None, mean_samples: int = 4096):
    """Returns (w [N, n_latent, D], losses).  ``noise``: explicit per-layer noise list (fixed over the optimisation, like
    ``randomize_noise=False`` runs) or None for fresh noise every step as in the reference."""
    if not target.is_cuda and getattr(generator, "requires_cuda", True):
        raise RuntimeError("input must be a CUDA tensor")
    with torch.no_grad():
        if mean_latent is None:
            mean_latent = generator.mean_latent(mean_samples)                      # [1, D]
    w = mean_latent.detach().reshape(1, 1, -1).repeat(target.shape[0], generator.n_latent, 1).clone().requires_grad_(True)
    if optim == "Adam":
        opt = torch.optim.Adam([w], lr=lr, betas=(0.5, 0.9))
    elif optim == "GD":
        opt = torch.optim.SGD([w], lr=lr, momentum=0.9)
    else:
        raise ValueError("optim must be 'Adam' or 'GD'")
    losses = []
    for _ in range(n_loops):
        out, _ = generator(w, input_is_latent=True, noise=noise)
        loss = reconstruction_loss(out, target)
        if extra_loss is not None:
            loss = loss + extra_loss(out, target)
        loss = loss.sum()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        losses.append(loss.detach())
    return w.detach(), torch.stack(losses)
