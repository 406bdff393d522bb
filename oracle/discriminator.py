"""Oracle restatement of the reference StyleGAN2 discriminator (test infrastructure only).

Follows reference graphs/stylegan_v2_real/networks.py: ``ConvLayer`` :517-565 (blur padding :531-535, stride-2
``EqualConv2d`` with ``scale = 1/sqrt(Cin k^2)`` :90-118, ``FusedLeakyReLU`` / ``ScaledLeakyReLU`` :164-173),
``ResBlock`` :568-588 (``(conv2(conv1(x)) + skip(x)) / sqrt(2)``), ``Discriminator`` :591-645 (minibatch standard
deviation :629-637, ``final_conv``, two ``EqualLinear`` :120-161).  It is a function of a rosinality-layout
``state_dict`` (keys ``convs.0.0.weight`` ...), works in the dtype of its input (float64 for arbiter runs) and uses
the oracle's own restatements of the two native ops (``oracle/ops.py``).

Pinning: ``tests/golden/ref_gpu_discriminator.npz`` holds logits and input gradients of the UNMODIFIED reference
module run on a B200 (``tests/golden/make_golden_ref_gpu_disc.py``); ``tests/test_oracle_discriminator.py`` checks
this file against it.
"""
from __future__ import annotations

import math

import torch
from torch.nn import functional as F

from .ops import fused_leaky_relu_ref, make_fir_kernel, upfirdn2d_ref


def _conv_layer_ref(sd, prefix, x, kernel_size, downsample=False, activate=True, bias=True, blur_taps=(1, 3, 3, 1)):
    i = 0
    if downsample:
        p = (len(blur_taps) - 2) + (kernel_size - 1)
        kern = sd.get(f"{prefix}.0.kernel")
        kern = make_fir_kernel(list(blur_taps), dtype=x.dtype) if kern is None else kern.to(x.dtype)
        x = upfirdn2d_ref(x, kern, pad=((p + 1) // 2, p // 2))
        i = 1
    w = sd[f"{prefix}.{i}.weight"].to(x.dtype)
    scale = 1.0 / math.sqrt(w.shape[1] * kernel_size ** 2)
    conv_bias = sd.get(f"{prefix}.{i}.bias") if (bias and not activate) else None
    x = F.conv2d(x, w * scale, bias=None if conv_bias is None else conv_bias.to(x.dtype), stride=2 if downsample else 1,
                 padding=0 if downsample else kernel_size // 2)
    if activate:
        if bias:
            x = fused_leaky_relu_ref(x, sd[f"{prefix}.{i + 1}.bias"].to(x.dtype))
        else:
            x = torch.where(x > 0, x, x * 0.2) * math.sqrt(2)
    return x


def _equal_linear_ref(sd, prefix, x, activation):
    w = sd[f"{prefix}.weight"].to(x.dtype)
    b = sd[f"{prefix}.bias"].to(x.dtype)
    y = x @ (w * (1.0 / math.sqrt(w.shape[1]))).t()
    if activation:
        return fused_leaky_relu_ref(y, b)
    return y + b


def discriminator_forward_ref(sd, image, size, stddev_group=4):
    """``sd``: state_dict (tensors), ``image`` [B, 3, size, size] -> logits [B, 1]."""
    log_size = int(math.log(size, 2))
    out = _conv_layer_ref(sd, "convs.0", image, 1)
    for j in range(1, log_size - 1):
        pre = f"convs.{j}"
        main = _conv_layer_ref(sd, pre + ".conv1", out, 3)
        main = _conv_layer_ref(sd, pre + ".conv2", main, 3, downsample=True)
        skip = _conv_layer_ref(sd, pre + ".skip", out, 1, downsample=True, activate=False, bias=False)
        out = (main + skip) / math.sqrt(2)
    b, c, h, w = out.shape
    group = min(b, stddev_group)
    s = out.reshape(group, b // group, 1, c, h, w)
    s = torch.sqrt(((s - s.mean(0, keepdim=True)) ** 2).mean(0) + 1e-8)       # biased variance over the group
    s = s.mean(dim=(2, 3, 4), keepdim=True).squeeze(2)                         # [b/group, 1, 1, 1]
    out = torch.cat([out, s.repeat(group, 1, h, w)], 1)
    out = _conv_layer_ref(sd, "final_conv", out, 3)
    out = out.reshape(b, -1)
    out = _equal_linear_ref(sd, "final_linear.0", out, True)
    return _equal_linear_ref(sd, "final_linear.1", out, False)
