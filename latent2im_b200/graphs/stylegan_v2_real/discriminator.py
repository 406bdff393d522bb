"""StyleGAN2 discriminator with the reference's checkpoint layout and call surface.

Drop-in for ``graphs/stylegan_v2_real/networks.py::Discriminator`` (reference networks.py:517-645) so that
``ckpt['d']`` loads key for key (``convs.0.0.weight``, ``convs.{i}.conv2.0.kernel``, ``final_linear.1.bias`` ...)
and ``TransformGraph.optimizeParametersAll`` can add the GAN term (reference transform_base.py:455-462).

Division of labour for this "next" row: the two ops the reference implements natively run on this repository's
sm_100a kernels through the C ABI -- the anti-aliasing blur in front of every stride-2 conv is ``l2i_upfirdn2d``
(reference Blur, networks.py:72-88) and every bias + leaky-relu is ``l2i_fused_bias_act`` (FusedLeakyReLU) -- while
the dense 3x3 / 1x1 convolutions stay on cuDNN through ``F.conv2d``, exactly where the reference has them
(networks.py:111-118).  The network is written as one functional pass over a flat parameter list; the two native
ops are arguments, which lets the CPU tests run the same wiring against the oracle's restatements of the ops.
"""
import math

import torch
from torch import nn
from torch.nn import functional as F

from .networks import _fir2d

# channels at each resolution for multiplier 1 (reference networks.py:580-590 uses 2 * this above 32 px)
_D_CHANNELS = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256, 128: 128, 256: 64, 512: 32, 1024: 16}


def _native_ops():
    from .op import fused_leaky_relu, upfirdn2d
    return upfirdn2d, fused_leaky_relu


def _blur_pad(kernel_size, n_taps=4, factor=2):
    """Padding of the blur in front of a stride-2 conv (reference ConvLayer, networks.py:531-535)."""
    p = (n_taps - factor) + (kernel_size - 1)
    return (p + 1) // 2, p // 2


class _ConvLayer(nn.Sequential):
    """Parameter holder in the reference's Sequential layout: [Blur]? + EqualConv2d + [FusedLeakyReLU]?
    (indices matter: they are the checkpoint key components)."""

    def __init__(self, in_channel, out_channel, kernel_size, downsample=False, blur_kernel=(1, 3, 3, 1), bias=True,
                 activate=True):
        mods = []
        if downsample:
            blur = nn.Module()
            blur.register_buffer("kernel", _fir2d(blur_kernel))
            mods.append(blur)
        conv = nn.Module()
        conv.weight = nn.Parameter(torch.randn(out_channel, in_channel, kernel_size, kernel_size))
        if bias and not activate:
            conv.bias = nn.Parameter(torch.zeros(out_channel))
        mods.append(conv)
        if activate and bias:
            act = nn.Module()
            act.bias = nn.Parameter(torch.zeros(out_channel))
            mods.append(act)
        super().__init__(*mods)
        self.kernel_size, self.downsample, self.activate, self.has_bias = kernel_size, downsample, activate, bias
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)

    def run(self, x, upfirdn2d, fused_leaky_relu):
        i = 0
        if self.downsample:
            x = upfirdn2d(x, self[0].kernel.to(x.dtype), pad=_blur_pad(self.kernel_size, self[0].kernel.shape[0]))
            i = 1
        conv = self[i]
        x = F.conv2d(x, conv.weight.to(x.dtype) * self.scale, bias=getattr(conv, "bias", None),
                     stride=2 if self.downsample else 1, padding=0 if self.downsample else self.kernel_size // 2)
        if self.activate:
            if self.has_bias:
                x = fused_leaky_relu(x, self[i + 1].bias.to(x.dtype))
            else:   # ScaledLeakyReLU (reference networks.py:164-173)
                x = F.leaky_relu(x, 0.2) * math.sqrt(2)
        return x

    def forward(self, x):
        return self.run(x, *_native_ops())


class _ResBlock(nn.Module):
    def __init__(self, in_channel, out_channel, blur_kernel=(1, 3, 3, 1)):
        super().__init__()
        self.conv1 = _ConvLayer(in_channel, in_channel, 3)
        self.conv2 = _ConvLayer(in_channel, out_channel, 3, downsample=True, blur_kernel=blur_kernel)
        self.skip = _ConvLayer(in_channel, out_channel, 1, downsample=True, blur_kernel=blur_kernel, activate=False, bias=False)

    def run(self, x, upfirdn2d, fused_leaky_relu):
        out = self.conv2.run(self.conv1.run(x, upfirdn2d, fused_leaky_relu), upfirdn2d, fused_leaky_relu)
        return (out + self.skip.run(x, upfirdn2d, fused_leaky_relu)) / math.sqrt(2)

    def forward(self, x):
        return self.run(x, *_native_ops())


class _EqualLinear(nn.Module):
    def __init__(self, in_dim, out_dim, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim))
        self.bias = nn.Parameter(torch.zeros(out_dim))
        self.activation = activation
        self.scale = 1 / math.sqrt(in_dim)

    def run(self, x, fused_leaky_relu):
        if self.activation:
            return fused_leaky_relu(F.linear(x, self.weight.to(x.dtype) * self.scale), self.bias.to(x.dtype))
        return F.linear(x, self.weight.to(x.dtype) * self.scale, bias=self.bias.to(x.dtype))


def minibatch_stddev(out, group_size=4):
    """One extra feature map holding the mean over (channel, pixel) of the per-group standard deviation
    (reference Discriminator.forward, networks.py:629-637, stddev_feat = 1)."""
    batch, channel, height, width = out.shape
    group = min(batch, group_size)
    s = out.view(group, -1, 1, channel, height, width)
    s = torch.sqrt(s.var(0, unbiased=False) + 1e-8)
    s = s.mean([2, 3, 4], keepdim=True).squeeze(2)
    return torch.cat([out, s.repeat(group, 1, height, width)], 1)


class Discriminator(nn.Module):
    def __init__(self, size, channel_multiplier=2, blur_kernel=(1, 3, 3, 1)):
        super().__init__()
        channels = {r: (c if r <= 32 else c * channel_multiplier) for r, c in _D_CHANNELS.items()}
        log_size = int(math.log(size, 2))
        convs = [_ConvLayer(3, channels[size], 1)]
        in_channel = channels[size]
        for i in range(log_size, 2, -1):
            out_channel = channels[2 ** (i - 1)]
            convs.append(_ResBlock(in_channel, out_channel, blur_kernel))
            in_channel = out_channel
        self.convs = nn.Sequential(*convs)
        self.stddev_group, self.stddev_feat = 4, 1
        self.final_conv = _ConvLayer(in_channel + 1, channels[4], 3)
        self.final_linear = nn.Sequential(_EqualLinear(channels[4] * 4 * 4, channels[4], activation="fused_lrelu"),
                                          _EqualLinear(channels[4], 1))
        self.size = size

    def run(self, image, upfirdn2d, fused_leaky_relu):
        out = image
        for layer in self.convs:
            out = layer.run(out, upfirdn2d, fused_leaky_relu)
        out = minibatch_stddev(out, self.stddev_group)
        out = self.final_conv.run(out, upfirdn2d, fused_leaky_relu)
        out = out.reshape(out.shape[0], -1)
        return self.final_linear[1].run(self.final_linear[0].run(out, fused_leaky_relu), fused_leaky_relu)

    def forward(self, image):
        """``image``: [B, 3, size, size] in about [-1, 1] -> logits [B, 1] (reference networks.py:626-645)."""
        return self.run(image, *_native_ops())
