"""Oracle restatement of the reference StyleGAN2 generator forward (test infrastructure only).

Functional: every routine takes the rosinality-format ``state_dict`` and
tensors, no ``nn.Module``.  The dense contractions themselves live in third-party code the
reference calls (``torch.nn.functional.conv2d / conv_transpose2d / linear`` -> on CPU: mkldnn /
native ATen; the reference pins only "PyTorch 1.6" in README.md:11, this image has torch 2.11);
the oracle calls the same ``torch.nn.functional`` entry points on CPU tensors, in float64 when it
acts as the arbiter.

Reference lines followed:
  PixelNorm + mapping MLP ........ networks.py:11-16, 129-156, 374-382
  ModulatedConv2d ................ networks.py:231-272
  Blur / Upsample ................ networks.py:30-48, 72-88, 197-203
  NoiseInjection ................. networks.py:275-286
  StyledConv / ToRGB ............. networks.py:330-358
  Generator.forward .............. networks.py:460-514 (latent / noise index maps)
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F

from .ops import fused_leaky_relu_ref, make_fir_kernel, upfirdn2d_ref


@dataclass
class GeneratorSpec:
    """Static shape facts of ``Generator(size, style_dim, n_mlp, channel_multiplier)``
    (networks.py:361-438)."""
    size: int
    style_dim: int = 512
    n_mlp: int = 8
    channel_multiplier: int = 2
    lr_mlp: float = 0.01
    blur_taps: tuple = (1, 3, 3, 1)

    @property
    def log_size(self) -> int:
        return int(math.log2(self.size))

    @property
    def num_layers(self) -> int:
        return (self.log_size - 2) * 2 + 1

    @property
    def n_latent(self) -> int:
        return self.log_size * 2 - 2

    def channels(self, res: int) -> int:
        cm = self.channel_multiplier
        table = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * cm, 128: 128 * cm,
                 256: 64 * cm, 512: 32 * cm, 1024: 16 * cm}
        return table[res]


def mapping_ref(sd, z, spec: GeneratorSpec):
    """``Generator.style``: PixelNorm then n_mlp x (EqualLinear + fused leaky relu)."""
    x = z * torch.rsqrt(torch.mean(z * z, dim=1, keepdim=True) + 1e-8)
    for i in range(1, spec.n_mlp + 1):
        w = sd[f"style.{i}.weight"].to(x.dtype)
        b = sd[f"style.{i}.bias"].to(x.dtype)
        scale = (1.0 / math.sqrt(w.shape[1])) * spec.lr_mlp
        x = F.linear(x, w * scale)
        x = fused_leaky_relu_ref(x, b * spec.lr_mlp)
    return x


def _modulation(sd, prefix, latent_i):
    w = sd[prefix + ".modulation.weight"].to(latent_i.dtype)
    b = sd[prefix + ".modulation.bias"].to(latent_i.dtype)
    return F.linear(latent_i, w * (1.0 / math.sqrt(w.shape[1])), bias=b)


def modulated_conv_ref(sd, prefix, x, latent_i, demodulate=True, upsample=False, blur_taps=(1, 3, 3, 1)):
    """``ModulatedConv2d.forward`` for the plain and upsample branches (the downsample branch is
    only used by the discriminator, out of scope)."""
    weight = sd[prefix + ".weight"].to(x.dtype)           # [1, Cout, Cin, k, k]
    _, cout, cin, k, _ = weight.shape
    batch, _, h, w = x.shape
    style = _modulation(sd, prefix, latent_i).view(batch, 1, cin, 1, 1)
    wmod = (1.0 / math.sqrt(cin * k * k)) * weight * style
    if demodulate:
        demod = torch.rsqrt(wmod.pow(2).sum([2, 3, 4]) + 1e-8)
        wmod = wmod * demod.view(batch, cout, 1, 1, 1)
    if upsample:
        wt = wmod.transpose(1, 2).reshape(batch * cin, cout, k, k)
        out = F.conv_transpose2d(x.reshape(1, batch * cin, h, w), wt, padding=0, stride=2, groups=batch)
        out = out.view(batch, cout, out.shape[2], out.shape[3])
        factor = 2
        p = (len(blur_taps) - factor) - (k - 1)
        pad = ((p + 1) // 2 + factor - 1, p // 2 + 1)
        fir = make_fir_kernel(blur_taps, gain=factor ** 2, dtype=x.dtype)
        out = upfirdn2d_ref(out, fir, pad=pad)
    else:
        out = F.conv2d(x.reshape(1, batch * cin, h, w), wmod.view(batch * cout, cin, k, k),
                       padding=k // 2, groups=batch)
        out = out.view(batch, cout, out.shape[2], out.shape[3])
    return out


def styled_conv_ref(sd, prefix, x, latent_i, noise, upsample=False, blur_taps=(1, 3, 3, 1)):
    """``StyledConv.forward``: modulated conv -> + weight*noise -> bias + leaky relu * sqrt(2).
    ``noise`` must be given (the oracle never draws random numbers itself)."""
    out = modulated_conv_ref(sd, prefix + ".conv", x, latent_i, True, upsample, blur_taps)
    out = out + sd[prefix + ".noise.weight"].to(x.dtype) * noise.to(x.dtype)
    return fused_leaky_relu_ref(out, sd[prefix + ".activate.bias"].to(x.dtype))


def to_rgb_ref(sd, prefix, x, latent_i, skip=None, blur_taps=(1, 3, 3, 1)):
    """``ToRGB.forward``: 1x1 modulated conv without demodulation + bias (+ upsampled skip)."""
    out = modulated_conv_ref(sd, prefix + ".conv", x, latent_i, demodulate=False)
    out = out + sd[prefix + ".bias"].to(x.dtype)
    if skip is not None:
        factor = 2
        fir = make_fir_kernel(blur_taps, gain=factor ** 2, dtype=x.dtype)
        p = fir.shape[0] - factor
        out = out + upfirdn2d_ref(skip, fir, up=factor, down=1, pad=((p + 1) // 2 + factor - 1, p // 2))
    return out


def generator_forward_ref(sd, latent, noise, spec: GeneratorSpec, return_intermediates=False):
    """``Generator.forward(styles=latent[B, n_latent, D], input_is_latent=True, noise=[...])``.

    ``noise`` is the explicit list of ``num_layers`` tensors, each ``[B or 1, 1, H, W]``.
    Latent index map: conv1<-0, to_rgb1<-1, block k: up-conv<-2k+1, conv<-2k+2, to_rgb<-2k+3.
    Noise index map: conv1<-0, block k: up-conv<-2k+1, conv<-2k+2."""
    dtype = latent.dtype
    batch = latent.shape[0]
    taps = spec.blur_taps
    inter = {}
    out = sd["input.input"].to(dtype).repeat(batch, 1, 1, 1)
    out = styled_conv_ref(sd, "conv1", out, latent[:, 0], noise[0], False, taps)
    skip = to_rgb_ref(sd, "to_rgb1", out, latent[:, 1], None, taps)
    inter["conv1"], inter["to_rgb1"] = out, skip
    i = 1
    for k in range(spec.log_size - 2):
        out = styled_conv_ref(sd, f"convs.{2 * k}", out, latent[:, i], noise[2 * k + 1], True, taps)
        inter[f"convs.{2 * k}"] = out
        out = styled_conv_ref(sd, f"convs.{2 * k + 1}", out, latent[:, i + 1], noise[2 * k + 2], False, taps)
        inter[f"convs.{2 * k + 1}"] = out
        skip = to_rgb_ref(sd, f"to_rgbs.{k}", out, latent[:, i + 2], skip, taps)
        inter[f"to_rgbs.{k}"] = skip
        i += 2
    return (skip, inter) if return_intermediates else skip


def clip_to_uint8_ref(img):
    """transform_base.py:551-552, 625-626: ``np.uint8(np.clip((x + 1) / 2 * 255, 0, 255))`` -
    a truncating cast, evaluated in the image's own precision."""
    return torch.clamp((img + 1) / 2.0 * 255, 0, 255).to(torch.uint8)
