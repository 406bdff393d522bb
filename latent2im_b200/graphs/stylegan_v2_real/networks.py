"""StyleGAN2 generator with the reference's Python surface, executed by the sm_100a C-ABI library.

Drop-in for ``graphs/stylegan_v2_real/networks.py::Generator`` of KelestZ/Latent2im (constructor
networks.py:361-369, ``forward`` :460-514, ``make_noise`` :440-447, ``mean_latent`` :449-455,
``get_latent`` :457-458).  The module tree exists so that rosinality-format checkpoints
(``ckpt['g_ema']``) load key for key; the sub-modules only own parameters.
All arithmetic of ``style`` and ``forward`` happens in ``libl2i_b200.so``:

* mapping network  -> ``l2i_generator_mapping``  (PixelNorm + n_mlp fused linear/bias/lrelu kernels)
* synthesis        -> ``l2i_generator_forward``  (implicit-GEMM modulated convs with fused
  demodulation / noise / bias / leaky-relu / next-layer modulation / ToRGB epilogues, fused
  blur kernel after the stride-2 transposed convs, fused skip up-sampling)

Noise: with ``noise=None, randomize_noise=True`` the per-layer noise tensors are drawn with
``torch.empty(B, 1, H, W).normal_()`` in the reference's execution order (conv1, convs[0], ...), in
float32 on the generator's device, so a seeded run consumes the CUDA Philox stream exactly like
the reference (NoiseInjection.forward, networks.py:281-286).
"""
import ctypes as C
import math
import os
import weakref

import torch
from torch import nn

from latent2im_b200 import _native as nt

_CHANNEL_TABLE = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256, 128: 128, 256: 64, 512: 32, 1024: 16}


def _fir2d(taps, gain=1.0):
    k = torch.tensor(taps, dtype=torch.float32)
    if k.ndim == 1:
        k = torch.outer(k, k)
    return k / k.sum() * gain


class _ParamOnly(nn.Module):
    """Sub-modules of the generator hold parameters in the checkpoint layout; they are executed
    only as part of ``Generator.forward`` (the fused native path has no per-module entry)."""

    def forward(self, *args, **kwargs):
        raise RuntimeError(
            f"{type(self).__name__} is executed inside Generator.forward by the native library; "
            "call the Generator (or latent2im_b200 ops) instead of the sub-module")


class PixelNorm(_ParamOnly):
    pass


class EqualLinear(_ParamOnly):
    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, lr_mul=1, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim) / lr_mul)
        self.bias = nn.Parameter(torch.full((out_dim,), float(bias_init))) if bias else None
        self.activation, self.lr_mul = activation, lr_mul
        self.scale = lr_mul / math.sqrt(in_dim)


class Blur(_ParamOnly):
    def __init__(self, kernel, pad, upsample_factor=1):
        super().__init__()
        self.register_buffer("kernel", _fir2d(kernel, float(upsample_factor ** 2) if upsample_factor > 1 else 1.0))
        self.pad = pad


class Upsample(_ParamOnly):
    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        self.register_buffer("kernel", _fir2d(kernel, float(factor ** 2)))
        p = self.kernel.shape[0] - factor
        self.pad = ((p + 1) // 2 + factor - 1, p // 2)


class ModulatedConv2d(_ParamOnly):
    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True, upsample=False,
                 blur_kernel=(1, 3, 3, 1)):
        super().__init__()
        self.in_channel, self.out_channel, self.kernel_size = in_channel, out_channel, kernel_size
        self.demodulate, self.upsample = demodulate, upsample
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        if upsample:
            p = (len(blur_kernel) - 2) - (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2 + 1, p // 2 + 1), upsample_factor=2)
        self.weight = nn.Parameter(torch.randn(1, out_channel, in_channel, kernel_size, kernel_size))
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1)


class NoiseInjection(_ParamOnly):
    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))


class ConstantInput(_ParamOnly):
    def __init__(self, channel, size=4):
        super().__init__()
        self.input = nn.Parameter(torch.randn(1, channel, size, size))


class _Bias(_ParamOnly):
    """Stands in for FusedLeakyReLU as a parameter holder (key ``activate.bias``)."""

    def __init__(self, channel):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))


class StyledConv(_ParamOnly):
    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=(1, 3, 3, 1)):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample,
                                    blur_kernel=blur_kernel)
        self.noise = NoiseInjection()
        self.activate = _Bias(out_channel)


class ToRGB(_ParamOnly):
    def __init__(self, in_channel, style_dim, upsample=True, blur_kernel=(1, 3, 3, 1)):
        super().__init__()
        if upsample:
            self.upsample = Upsample(blur_kernel)
        self.conv = ModulatedConv2d(in_channel, 3, 1, style_dim, demodulate=False)
        self.bias = nn.Parameter(torch.zeros(1, 3, 1, 1))


class _Mapping(nn.Sequential):
    """``Generator.style``: a Sequential in the checkpoint layout (keys ``style.{1..n}.*``) whose call
    runs the native mapping kernels of the owning generator (TransformGraph.get_w calls
    ``netG.style(z)``, transform_base.py:372-373)."""

    def bind(self, owner):
        object.__setattr__(self, "_owner", weakref.ref(owner))
        return self

    def forward(self, z):
        return self._owner().get_latent(z)


class _NativeHandle:
    """Owns one ``l2i_generator_t`` (weights packed for one device / dtype / max batch)."""

    def __init__(self, gen, device, dtype, max_batch):
        lib = nt.load()
        self.lib, self.device, self.dtype, self.max_batch = lib, device, dtype, max_batch
        self.handle = C.c_void_p()
        taps = (C.c_float * len(gen.blur_kernel))(*[float(t) for t in gen.blur_kernel])
        with torch.cuda.device(device):
            nt.check(lib.l2i_generator_create(C.byref(self.handle), gen.size, gen.style_dim, gen.n_mlp,
                                              gen.channel_multiplier, taps, len(gen.blur_kernel), float(gen.lr_mlp),
                                              nt.dtype_code(dtype), max_batch), "generator_create")
        self.version = None
        self.training = False

    def upload(self, gen):
        # (data_ptr, version counter) of every parameter, read through a cached list of (module, name) slots: nn.Module.parameters()
        # walks the module tree with de-duplication sets and cost 0.3 ms per call - a third of the host time of a 256-px batch-4
        # forward (tools/probes/host_overhead.py).  In-place updates, load_state_dict, .to() and assigning a new Parameter to an
        # existing module are all seen; replacing a whole sub-module is not (call gen.invalidate_native() after such surgery).
        version = tuple((p.data_ptr(), p._version) for p in gen._param_list())
        if version == self.version:
            return
        with torch.cuda.device(self.device):
            st = nt.stream_ptr(self.device)
            for key, t in gen.state_dict().items():
                if key.startswith("noises.") or key.endswith(".kernel"):
                    continue
                t32 = t.detach().to(device=self.device, dtype=torch.float32).contiguous()
                nt.check(self.lib.l2i_generator_set_param(self.handle, key.encode(), t32.data_ptr(), t32.numel(), st),
                         f"generator_set_param({key})")
            nt.check(self.lib.l2i_generator_finalize(self.handle, st), "generator_finalize")
            torch.cuda.current_stream(self.device).synchronize()  # staging copies above may be freed now
        self.version = version

    def __del__(self):
        try:
            if self.handle:
                self.lib.l2i_generator_destroy(self.handle)
        except Exception:
            pass


class _SynthesisFn(torch.autograd.Function):
    """Differentiable synthesis: forward in training mode (activations kept natively), backward =
    ``l2i_generator_backward`` (data gradient w.r.t. the W+ latent only; the generator is frozen)."""

    @staticmethod
    def forward(ctx, latent, gen, noise):
        image = gen._synthesize(latent, noise=noise, training=True)
        ctx.gen = gen
        ctx.token = gen._last_train
        ctx.lat_shape = latent.shape
        return image

    @staticmethod
    def backward(ctx, grad_image):
        gen = ctx.gen
        if gen._last_train is not ctx.token:
            raise RuntimeError("Generator backward: another forward (training or inference) ran on this generator before this "
                               "backward; the native library keeps the activations and style tables of one forward at a time")
        h, batch, _ = ctx.token
        g = grad_image.contiguous().float()
        grad_latent = torch.empty(ctx.lat_shape, device=g.device, dtype=torch.float32)
        with torch.cuda.device(g.device):
            nt.check(h.lib.l2i_generator_backward(h.handle, grad_latent.data_ptr(), g.data_ptr(), batch, nt.stream_ptr(g.device)),
                     "generator_backward")
        return grad_latent, None, None


class Generator(nn.Module):
    def __init__(self, size, style_dim, n_mlp, channel_multiplier=2, blur_kernel=[1, 3, 3, 1], lr_mlp=0.01):
        super().__init__()
        self.size, self.style_dim, self.n_mlp = size, style_dim, n_mlp
        self.channel_multiplier, self.blur_kernel, self.lr_mlp = channel_multiplier, list(blur_kernel), lr_mlp
        self.style = _Mapping(PixelNorm(), *[
            EqualLinear(style_dim, style_dim, lr_mul=lr_mlp, activation="fused_lrelu") for _ in range(n_mlp)]).bind(self)
        self.channels = {r: (c if r <= 32 else c * channel_multiplier) for r, c in _CHANNEL_TABLE.items()}
        self.log_size = int(math.log(size, 2))
        self.num_layers = (self.log_size - 2) * 2 + 1
        self.n_latent = self.log_size * 2 - 2

        self.input = ConstantInput(self.channels[4])
        self.conv1 = StyledConv(self.channels[4], self.channels[4], 3, style_dim, blur_kernel=blur_kernel)
        self.to_rgb1 = ToRGB(self.channels[4], style_dim, upsample=False)
        self.convs, self.upsamples, self.to_rgbs = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        self.noises = nn.Module()
        for i in range(self.num_layers):
            r = 2 ** ((i + 5) // 2)
            self.noises.register_buffer(f"noise_{i}", torch.randn(1, 1, r, r))
        cin = self.channels[4]
        for i in range(3, self.log_size + 1):
            cout = self.channels[2 ** i]
            self.convs.append(StyledConv(cin, cout, 3, style_dim, upsample=True, blur_kernel=blur_kernel))
            self.convs.append(StyledConv(cout, cout, 3, style_dim, blur_kernel=blur_kernel))
            self.to_rgbs.append(ToRGB(cout, style_dim))
            cin = cout

        self._native = {}
        env = os.environ.get("L2I_DTYPE", "bf16").lower()
        self.compute_dtype = torch.float32 if env in ("fp32", "f32", "float32") else torch.bfloat16
        self.max_batch = 0

    # ---- native plumbing ------------------------------------------------------------------------
    def set_native(self, dtype=None, max_batch=None):
        """Selects the arithmetic of the synthesis kernels (``torch.bfloat16``: tcgen05 tensor-core
        path, fp32 accumulation; ``torch.float32``: CUDA-core arbiter path) and pre-sizes the
        workspace for ``max_batch`` samples."""
        if dtype is not None:
            if dtype not in (torch.float32, torch.bfloat16):
                raise ValueError("compute dtype must be torch.float32 or torch.bfloat16")
            self.compute_dtype = dtype
        if max_batch is not None:
            self.max_batch = max(self.max_batch, int(max_batch))
        return self

    def _param_list(self):
        slots = self.__dict__.get("_param_slots")
        if slots is None:
            slots = [(m, n) for m in self.modules() for n in m._parameters]
            object.__setattr__(self, "_param_slots", slots)
        return [m._parameters[n] for m, n in slots if m._parameters.get(n) is not None]

    def invalidate_native(self):
        """Forget the cached parameter slots (after replacing sub-modules); the next forward re-checks every weight."""
        self.__dict__.pop("_param_slots", None)
        for h in self._native.values():
            h.version = None

    def _handle(self, device, batch):
        if device.type != "cuda":
            raise RuntimeError("input must be a CUDA tensor")  # reference ops: TORCH_CHECK(is_cuda)
        key = (device.index if device.index is not None else torch.cuda.current_device(), self.compute_dtype)
        h = self._native.get(key)
        need = max(batch, self.max_batch, 1)
        if h is None or h.max_batch < need:
            self._native.pop(key, None)
            h = None
            h = _NativeHandle(self, torch.device("cuda", key[0]), self.compute_dtype, need)
            self._native[key] = h
        h.upload(self)
        return h

    def _device(self):
        return self.input.input.device

    # ---- reference API --------------------------------------------------------------------------
    def make_noise(self):
        device = self._device()
        noises = [torch.randn(1, 1, 4, 4, device=device)]
        for i in range(3, self.log_size + 1):
            noises += [torch.randn(1, 1, 2 ** i, 2 ** i, device=device) for _ in range(2)]
        return noises

    def get_latent(self, input):
        nt.require_cuda(input, "input")
        z = input.detach().to(torch.float32).contiguous()
        lead = z.shape[:-1]
        z2 = z.reshape(-1, self.style_dim)
        h = self._handle(z.device, z2.shape[0])
        w = torch.empty_like(z2)
        with torch.cuda.device(z.device):
            nt.check(h.lib.l2i_generator_mapping(h.handle, nt.ptr(w), nt.ptr(z2), z2.shape[0], nt.stream_ptr(z.device)),
                     "generator_mapping")
        return w.reshape(*lead, self.style_dim)

    def mean_latent(self, n_latent):
        latent_in = torch.randn(n_latent, self.style_dim, device=self._device())
        return self.get_latent(latent_in).mean(0, keepdim=True)

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_native"] = {}
        state.pop("_param_slots", None)
        state.pop("_last", None)
        state.pop("_last_train", None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self.style.bind(self)

    def forward(self, styles, return_latents=False, inject_index=None, truncation=1, truncation_latent=None,
                input_is_latent=False, noise=None, randomize_noise=True):
        if input_is_latent:
            latent = styles
        else:
            # the reference leaves ``latent`` undefined on this branch (networks.py:471-494, NameError);
            # this implements the upstream rosinality semantics it was derived from
            ws = [self.get_latent(s) for s in styles]
            if truncation < 1:
                ws = [truncation_latent + truncation * (w - truncation_latent) for w in ws]
            if len(ws) < 2:
                latent = ws[0]
                if latent.ndim < 3:
                    latent = latent.unsqueeze(1).expand(-1, self.n_latent, -1)
            else:
                import random
                if inject_index is None:
                    inject_index = random.randint(1, self.n_latent - 1)
                latent = torch.cat([ws[0].unsqueeze(1).expand(-1, inject_index, -1),
                                    ws[1].unsqueeze(1).expand(-1, self.n_latent - inject_index, -1)], 1)
        if latent.ndim != 3 or latent.shape[1] != self.n_latent or latent.shape[2] != self.style_dim:
            raise RuntimeError(f"latent must have shape [B, {self.n_latent}, {self.style_dim}], got {tuple(latent.shape)}")
        nt.require_cuda(latent, "input")
        image = self.synthesize(latent, noise=noise, randomize_noise=randomize_noise)
        return (image, latent) if return_latents else (image, None)

    def _noise_list(self, batch, device, noise, randomize_noise):
        if noise is None:
            if randomize_noise:
                noise = [torch.empty(batch, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2), device=device,
                                     dtype=torch.float32).normal_() for i in range(self.num_layers)]
            else:
                noise = [getattr(self.noises, f"noise_{i}") for i in range(self.num_layers)]
        if len(noise) != self.num_layers:
            raise RuntimeError(f"noise must list {self.num_layers} tensors")
        out = []
        for i, n in enumerate(noise):
            if n is None:  # per-layer None inside a list: fresh noise, like NoiseInjection
                r = 2 ** ((i + 5) // 2)
                n = torch.empty(batch, 1, r, r, device=device, dtype=torch.float32).normal_()
            r = 2 ** ((i + 5) // 2)
            if n.shape[-2:] != (r, r) or n.shape[0] not in (1, batch):
                raise RuntimeError(f"noise[{i}] must be [1 or {batch}, 1, {r}, {r}], got {tuple(n.shape)}")
            out.append(n.detach().to(device=device, dtype=torch.float32).contiguous())
        return out

    def synthesize(self, latent, noise=None, randomize_noise=True, want_uint8=False, want_float=True, out_uint8=None):
        if torch.is_grad_enabled() and latent.requires_grad:
            if want_uint8:
                raise RuntimeError("the uint8 image is not differentiable; call synthesize under torch.no_grad()")
            nz = self._noise_list(latent.shape[0], latent.device, noise, randomize_noise)
            return _SynthesisFn.apply(latent, self, nz)
        return self._synthesize(latent, noise, randomize_noise, want_uint8, want_float, training=False, out_uint8=out_uint8)

    def _synthesize(self, latent, noise=None, randomize_noise=True, want_uint8=False, want_float=True, training=False,
                    out_uint8=None):
        """``latent`` [B, n_latent, D] -> image [B, 3, size, size] float32 (and / or the uint8 NHWC
        image ``clip((x + 1) / 2 * 255)`` the reference computes on the host, transform_base.py:625-626)."""
        device = latent.device
        batch = latent.shape[0]
        h = self._handle(device, batch)
        lat = latent.detach()
        if lat.dtype != torch.float32:
            lat = lat.float()
        if lat.stride(2) != 1:
            lat = lat.contiguous()
        nz = self._noise_list(batch, device, noise, randomize_noise)
        nz_ptrs = (C.c_void_p * self.num_layers)(*[n.data_ptr() for n in nz])
        nz_batch = (C.c_int * self.num_layers)(*[n.shape[0] for n in nz])
        image = torch.empty(batch, 3, self.size, self.size, device=device, dtype=torch.float32) if want_float else None
        image_u8 = None
        if want_uint8:
            image_u8 = out_uint8 if out_uint8 is not None else torch.empty(batch, self.size, self.size, 3, device=device, dtype=torch.uint8)
            if image_u8.shape != (batch, self.size, self.size, 3) or image_u8.dtype != torch.uint8 or not image_u8.is_contiguous():
                raise RuntimeError("out_uint8 must be a contiguous uint8 tensor of shape [B, size, size, 3]")
        with torch.cuda.device(device):
            if training or h.training:
                nt.check(h.lib.l2i_generator_set_training(h.handle, 1 if training else 0), "generator_set_training")
                h.training = training
            nt.check(h.lib.l2i_generator_forward(h.handle, lat.data_ptr(), lat.stride(0), lat.stride(1), nz_ptrs, nz_batch,
                                                 nt.ptr(image), nt.ptr(image_u8), batch, nt.stream_ptr(device)),
                     "generator_forward")
        if training:
            self._last_train = (h, batch, nz)  # the noise tensors must outlive the backward pass
        elif getattr(self, "_last_train", None) is not None and self._last_train[0] is h:
            # an inference forward on the SAME native handle rewrote its style / demod tables and activation buffers: the
            # pending backward of the earlier training forward would silently mix them with its saved activations
            self._last_train = None
        self._last = (h, batch, nz)  # keeps the noise tensors alive until the next call
        if want_uint8 and want_float:
            return image, image_u8
        return image_u8 if want_uint8 else image

    def read_activation(self, name):
        """Debug tap: fp32 NCHW copy of an internal activation of the last forward."""
        h, batch, _ = self._last
        if name.startswith("skip."):
            k = int(name[5:])
            res, ch = 4 * 2 ** k, 3
        else:
            mod = self.conv1 if name == "conv1" else self.convs[int(name.split(".")[1])]
            ch = mod.conv.out_channel
            res = 4 if name == "conv1" else 2 ** (3 + int(name.split(".")[1]) // 2)
        out = torch.empty(batch, ch, res, res, device=h.device, dtype=torch.float32)
        with torch.cuda.device(h.device):
            nt.check(h.lib.l2i_generator_read_activation(h.handle, name.encode(), out.data_ptr(), out.numel(), batch,
                                                         nt.stream_ptr(h.device)), "generator_read_activation")
        return out


# the discriminator lives in its own module; re-exported under the reference's names
from .discriminator import Discriminator, _ConvLayer as ConvLayer, _ResBlock as ResBlock  # noqa: E402,F401
