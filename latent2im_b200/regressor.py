"""Inference-time preparation of the (frozen, eval-mode) attribute regressor (transform_base.py:396-403, 522-534).

The regressor is a stock torchvision ResNet-50 and stays on cuDNN.  At the full-resolution
inputs of the walk-training step it is bound by elementwise traffic, not by its convolutions: every BatchNorm is a separate
read + write of the activation.  Because the regressor is frozen and in eval mode, each BatchNorm is an affine map with
constant coefficients and folds exactly into the preceding convolution's weights and bias (fp32 fold, done once)."""
from __future__ import annotations

import copy

import torch
from torch.nn.utils.fusion import fuse_conv_bn_eval


def fold_batchnorm(model: torch.nn.Module, inplace: bool = False) -> torch.nn.Module:
    """Returns ``model`` (eval mode) with every ``Conv2d -> BatchNorm2d`` pair of sibling sub-modules replaced by one conv.
    Forward values are unchanged up to fp32 rounding; gradients w.r.t. the INPUT are unchanged as well (the map is the
    same affine function), which is all the walk-training step needs - the regressor's own parameters are frozen."""
    if model.training:
        raise RuntimeError("fold_batchnorm needs an eval-mode model (running statistics are folded)")
    if not inplace:
        model = copy.deepcopy(model)

    def rec(mod):
        prev = None
        for name in list(mod._modules.keys()):
            child = mod._modules[name]
            if isinstance(child, torch.nn.BatchNorm2d) and prev is not None and isinstance(mod._modules[prev], torch.nn.Conv2d):
                mod._modules[prev] = fuse_conv_bn_eval(mod._modules[prev], child)
                mod._modules[name] = torch.nn.Identity()
            else:
                rec(child)
            prev = name

    rec(model)
    return model
