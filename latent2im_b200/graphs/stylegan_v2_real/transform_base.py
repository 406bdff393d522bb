"""Walk modules and the TransformGraph orchestrator over the native hot path.

Drop-in for ``graphs/stylegan_v2_real/transform_base.py`` of KelestZ/Latent2im: the three W-space
walk modules on the path (``WalkLinearMultiW`` :140-165, ``WalkMlpMultiW`` :168-204,
``WalkNonLinearW`` :207-243) keep their class names, constructor signatures and parameter names
(``w``, ``linear.{0,2,4}.*``, ``embed.*``) so whole-module pickles written by the reference load;
``TransformGraph`` keeps the method set the entry scripts call.
The walk step runs in ``l2i_walk_linear_fwd`` / ``l2i_walk_combine`` + ``l2i_linear_fwd``; the
walk-parameter gradient in ``l2i_walk_linear_bwd``.

Deviations from the shipped reference, all documented in DESIGN.md:
* size / batch / dtype come from ``constants`` instead of being hard-coded to 256 px, batch 4;
* ``WalkMlpMultiW`` with ``layers=...`` applies the same MLP to the chosen layers (the reference
  calls ``self.linear(input[i], 1)``, a TypeError); ``WalkNonLinearW`` accepts the keyword call
  ``walk(ws, alpha=..., layers=...)`` that ``get_w_new_tensor`` makes (TypeError as shipped);
* G-forward #1 / R-forward #1 of the training step run without autograd (the branch contributes
  exactly zero to the walk gradient);
* the discriminator (own ops + cuDNN convs) and VGG19 (stock) loss terms are only built when their loss is requested.
"""
import os

import numpy as np
import torch
from torch import nn
from torch.autograd import Function

from latent2im_b200 import _native as nt

from . import constants


# ------------------------------------------------------------------------------------------------
# native walk steps as autograd Functions
# ------------------------------------------------------------------------------------------------
def _as_latent_block(ws):
    """list of n_latent [B, D] tensors -> (tensor, batch_stride, layer_stride).  The W+ list built
    by get_w is one tensor repeated (transform_base.py:372-378): no copy, layer stride 0."""
    first = ws[0]
    nt.require_cuda(first, "input")
    if all(w is first for w in ws):
        base = first.detach().float().contiguous()
        return base, base.stride(0), 0
    blk = torch.stack([w.detach().float() for w in ws], 1).contiguous()
    return blk, blk.stride(0), blk.stride(1)


def _layer_mask(n, layers):
    if layers is None:
        return (1 << n) - 1
    m = 0
    for i in layers:
        i = int(i)
        if 0 <= i < n:
            m |= 1 << i
    return m



def _missing_checkpoint(what, path):
    """The reference fails hard in torch.load; random weights are only for benchmarks / tests that ask for them."""
    if not getattr(constants, "allow_random_init", False):
        raise FileNotFoundError(f"{what} checkpoint not found: {path!r} (set L2I_ALLOW_RANDOM_INIT=1 / --allow_random_init "
                                f"to run on random-init weights)")
    import warnings
    warnings.warn(f"{what} checkpoint {path!r} not found: running on RANDOM-INIT weights (allow_random_init)", stacklevel=2)


def _report_load(result, path):
    """load_state_dict(strict=False) result: FIR-kernel / noise buffers may legitimately differ, anything else is loud."""
    missing = [k for k in result.missing_keys if not (k.endswith(".kernel") or k.startswith("noises."))]
    if missing or result.unexpected_keys:
        import warnings
        warnings.warn(f"{path}: missing keys {missing[:8]}{'...' if len(missing) > 8 else ''}, unexpected keys "
                      f"{list(result.unexpected_keys)[:8]}", stacklevel=2)

class _WalkLinearFn(Function):
    @staticmethod
    def forward(ctx, w_param, alpha, mask, n_latent, *ws):
        blk, bs, ls = _as_latent_block(ws)
        B, D = ws[0].shape
        al = alpha.detach().float().contiguous()
        wp = w_param.detach().float().contiguous()
        out = torch.empty(B, n_latent, D, device=blk.device, dtype=torch.float32)
        with torch.cuda.device(blk.device):
            nt.check(nt.load().l2i_walk_linear_fwd(out.data_ptr(), blk.data_ptr(), bs, ls, nt.ptr(al), nt.ptr(wp), B,
                                                   al.shape[1], n_latent, D, mask, nt.stream_ptr(blk.device)),
                     "walk_linear_fwd")
        ctx.save_for_backward(al)
        ctx.cfg = (mask, n_latent, D, wp.shape[0], len(ws), all(w is ws[0] for w in ws))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (al,) = ctx.saved_tensors
        mask, n_latent, D, A, n_in, shared = ctx.cfg
        g = grad_out.contiguous().float()
        gw = torch.empty(A, n_latent, D, device=g.device, dtype=torch.float32)
        with torch.cuda.device(g.device):
            nt.check(nt.load().l2i_walk_linear_bwd(gw.data_ptr(), g.data_ptr(), al.data_ptr(), g.shape[0], A, n_latent, D,
                                                   mask, nt.stream_ptr(g.device)), "walk_linear_bwd")
        # d out_i / d in_i = I; d out / d alpha is never needed on the training path (epsilon is a value)
        gins = tuple(g[:, i] for i in range(n_in))
        return (gw, None, None, None) + gins


class _LinearFn(Function):
    """y = act(x W^T + b) through l2i_linear_fwd; backward (gx, gW, gb, activation derivative) through l2i_linear_bwd."""

    @staticmethod
    def forward(ctx, x, W, b, act):
        nt.require_cuda(x, "input")
        x2 = x.detach().float().contiguous()
        Wc, bc = W.detach().float().contiguous(), b.detach().float().contiguous()
        B, K = x2.shape
        N = Wc.shape[0]
        y = torch.empty(B, N, device=x2.device, dtype=torch.float32)
        with torch.cuda.device(x2.device):
            nt.check(nt.load().l2i_linear_fwd(y.data_ptr(), N, x2.data_ptr(), K, Wc.data_ptr(), bc.data_ptr(), B, N, K, 1.0, 1.0,
                                              1 if act else 0, 0.2, 1.0, nt.stream_ptr(x2.device)), "linear_fwd")
        ctx.save_for_backward(x2, Wc, y)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, gy):
        x2, Wc, y = ctx.saved_tensors
        g = gy.contiguous().float()
        B, K = x2.shape
        N = Wc.shape[0]
        dev = g.device
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        gx = torch.empty(B, K, device=dev, dtype=torch.float32) if need_x else None
        gW = torch.empty(N, K, device=dev, dtype=torch.float32) if need_w else None
        gb = torch.empty(N, device=dev, dtype=torch.float32) if need_w else None
        with torch.cuda.device(dev):   # leaky-relu derivative (from the saved output's sign), gx, gW and gb in two native launches
            nt.check(nt.load().l2i_linear_bwd(nt.ptr(gx), nt.ptr(gW), nt.ptr(gb), g.data_ptr(), y.data_ptr(), x2.data_ptr(),
                                              Wc.data_ptr(), B, N, K, 1 if ctx.act else 0, 0.2, nt.stream_ptr(dev)), "linear_bwd")
        return gx, gW, gb, None


def _mlp(seq, x):
    mods = list(seq)
    for j, m in enumerate(mods):
        if isinstance(m, nn.Linear):
            act = j + 1 < len(mods) and isinstance(mods[j + 1], nn.LeakyReLU)
            x = _LinearFn.apply(x, m.weight, m.bias, act)
    return x


class _CombineFn(Function):
    """out_i = in_i + coef * d_i (or in_i + d_i / ||d_i||) through l2i_walk_combine / l2i_walk_combine_bwd: the MLP walks
    train through native kernels only (the reference differentiates torch ops, transform_base.py:188-193, 227-229)."""

    @staticmethod
    def forward(ctx, dblk, coef, mask, normalize, shared_d, n, *ws):
        blk, bs, ls = _as_latent_block(ws)
        B, D = ws[0].shape
        d = dblk.detach().float().contiguous()
        dbs, dls = (d.stride(0), 0) if shared_d else (d.stride(0), d.stride(1))
        c = coef.detach().float().reshape(-1).contiguous() if coef is not None else None
        out = torch.empty(B, n, D, device=blk.device, dtype=torch.float32)
        with torch.cuda.device(blk.device):
            nt.check(nt.load().l2i_walk_combine(out.data_ptr(), blk.data_ptr(), bs, ls, d.data_ptr(), dbs, dls, nt.ptr(c), B, n, D,
                                                mask, 1 if normalize else 0, nt.stream_ptr(blk.device)), "walk_combine")
        ctx.save_for_backward(d, c) if c is not None else ctx.save_for_backward(d)
        ctx.cfg = (mask, normalize, shared_d, n, D, dbs, dls, c is not None, len(ws))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        mask, normalize, shared_d, n, D, dbs, dls, has_c, n_in = ctx.cfg
        saved = ctx.saved_tensors
        d, c = saved[0], (saved[1] if has_c else None)
        g = grad_out.contiguous().float()
        B = g.shape[0]
        gd = None
        if ctx.needs_input_grad[0]:
            gd = torch.empty_like(d)
            scratch = torch.empty(B, n, D, device=g.device, dtype=torch.float32) if shared_d else None
            with torch.cuda.device(g.device):
                nt.check(nt.load().l2i_walk_combine_bwd(gd.data_ptr(), nt.ptr(scratch), g.data_ptr(), d.data_ptr(), dbs, dls,
                                                        nt.ptr(c), B, n, D, mask, 1 if normalize else 0,
                                                        nt.stream_ptr(g.device)), "walk_combine_bwd")
        # d out_i / d in_i = I; the step size (alpha) is a value on the training path, never a parameter
        gins = tuple(g[:, i] if ctx.needs_input_grad[6 + i] else None for i in range(n_in))
        return (gd, None, None, None, None, None) + gins


def _combine(ws, d_list, coef, mask, normalize):
    """out_i = in_i + coef * d_i (or in_i + d_i/||d_i||) for the layers in `mask`: one native launch forward, one or two
    backward, with or without grad."""
    n = len(ws)
    fill = next((t for t in d_list if t is not None), None)
    if fill is None:
        return list(ws)
    d_list = [fill if t is None else t for t in d_list]
    shared = all(t is d_list[0] for t in d_list)
    dblk = d_list[0] if shared else torch.stack(list(d_list), 1)
    out = _CombineFn.apply(dblk, coef, mask, bool(normalize), shared, n, *ws)
    return list(out.unbind(1))


# ------------------------------------------------------------------------------------------------
# walk modules (class / parameter names as in the reference)
# ------------------------------------------------------------------------------------------------
class WalkLinearMultiW(nn.Module):
    """``out_i = in_i + alpha @ w[:, i, :]`` for every (or the selected) W+ layer."""

    def __init__(self, dim_z, step, Nsliders, attrList):
        super().__init__()
        self.dim_z, self.step = dim_z, step
        n_latent = (step + 1) * 2
        # numpy global RNG, as the reference (transform_base.py:147)
        self.w = nn.Parameter(torch.Tensor(np.random.normal(0.0, 0.02, [len(attrList), n_latent, dim_z])))

    def forward(self, input, alpha, layers=None, name=None, index_=None):
        n = len(input)
        alpha = alpha.to(input[0].device)
        # the kernel indexes w as [A][n][D] from raw pointers: a walk trained for another resolution (n_latent) or attribute
        # count must fail here, like the reference's `self.w[:, i, :]` / `alpha @ w` would (IndexError / shape error)
        if alpha.ndim != 2 or tuple(self.w.shape) != (alpha.shape[1], n, input[0].shape[1]):
            raise RuntimeError(f"WalkLinearMultiW: walk parameter {tuple(self.w.shape)} does not match alpha {tuple(alpha.shape)}, "
                               f"{n} latent layers of width {input[0].shape[1]} (expected w = [A, n_latent, D])")
        out = _WalkLinearFn.apply(self.w, alpha, _layer_mask(n, layers), n, *input)
        return list(out.unbind(1))


class WalkMlpMultiW(nn.Module):
    """``out_i = in_i + alpha[:, 0:1] * MLP(in_i)``, MLP = D -> 2D -> 2D -> D with LeakyReLU(0.2)."""

    def __init__(self, dim_z, step, Nsliders, attrList):
        super().__init__()
        self.dim_z, self.step, self.Nsliders = dim_z, step, Nsliders
        self.linear = nn.Sequential(nn.Linear(dim_z, 2 * dim_z), nn.LeakyReLU(0.2, True),
                                    nn.Linear(2 * dim_z, 2 * dim_z), nn.LeakyReLU(0.2, True),
                                    nn.Linear(2 * dim_z, dim_z))

    def forward(self, input, alpha, layers=None, name=None, index_=None):
        n = len(input)
        al = alpha[:, 0:1].to(input[0].device)
        mask = _layer_mask(n, layers)
        cache, d = {}, []
        for i in range(n):  # one MLP evaluation per distinct input tensor (the W+ list is usually one tensor)
            if not (mask >> i) & 1:
                d.append(None)
                continue
            key = id(input[i])
            if key not in cache:
                cache[key] = _mlp(self.linear, input[i])
            d.append(cache[key])
        return _combine(list(input), d, al, mask, normalize=False)


class WalkNonLinearW(nn.Module):
    """``e = embed(alpha[:, 0:1] x10)``; ``d = MLP([e, in_i])``; ``out_i = in_i + d/||d||``."""

    def __init__(self, dim_z, step, Nsliders, attrList):
        super().__init__()
        self.dim_z, self.step, self.Nsliders = dim_z, step, Nsliders
        self.embed = nn.Linear(10, dim_z // 2)
        self.linear = nn.Sequential(nn.Linear(dim_z // 2 + dim_z, 2 * dim_z), nn.LeakyReLU(0.2, True),
                                    nn.Linear(2 * dim_z, dim_z))

    def forward(self, input, name=None, alpha=None, index_=None, layers=None):
        if alpha is None and torch.is_tensor(name):  # keyword-less call walk(ws, alpha)
            name, alpha = None, name
        n = len(input)
        al = alpha[:, 0:1].to(input[0].device)
        e = _LinearFn.apply(al.repeat(1, 10), self.embed.weight, self.embed.bias, False)
        mask = _layer_mask(n, layers)
        cache, d = {}, []
        for i in range(n):
            if not (mask >> i) & 1:
                d.append(None)
                continue
            key = id(input[i])
            if key not in cache:
                cache[key] = _mlp(self.linear, torch.cat([e, input[i].float()], 1))
            d.append(cache[key])
        return _combine(list(input), d, None, mask, normalize=layers is None)


# ------------------------------------------------------------------------------------------------
# orchestrator
# ------------------------------------------------------------------------------------------------
class _Module:
    """Stands in for stylegan2.StyleGAN (stylegan2.py:19-64): holder of netG (and netD when built)."""

    def __init__(self):
        self.netG = None
        self.netD = None


class TransformGraph:
    def __init__(self, lr, walk_type, nsliders, loss_type, eps, N_f, trainEmbed, attrList, attrTable, layers,
                 stylegan_opts):
        assert loss_type in ["l2", "lpips"], "unimplemented loss"
        self.lr = lr
        self.useGPU = constants.useGPU
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.img_size = constants.resolution
        self.dim_z = constants.DIM_Z
        self.BATCH_SIZE = constants.BATCH_SIZE
        self.num_channels = constants.NUM_CHANNELS
        self.module = self.get_stylegan2_module()
        self.regressor, self.reg_optmizer = self.get_reg_module()
        self.attrTable, self.attrList = attrTable, attrList
        self.attrIdx = self.get_attr_idx()
        self.Nsliders = nsliders
        self.trainEmbed = trainEmbed
        self.stylegan_opts = stylegan_opts
        self.layers = layers
        self.walk_type = walk_type
        # (step + 1) * 2 == n_latent; the reference hard-codes step = 6 for 256 px (:285)
        self.step = self.module.netG.log_size - 2
        self.is_mlp = bool(getattr(constants, "walk_is_mlp", False))
        latent = getattr(stylegan_opts, "latent", "w")
        if walk_type == "linear":
            if trainEmbed:
                raise NotImplementedError("WalkEmbed is declared unused by the reference (transform_base.py:24-27)")
            if latent != "w":
                raise NotImplementedError("Not implemented setting of linear transformation for z")
            cls = WalkMlpMultiW if self.is_mlp else WalkLinearMultiW
            self.walk = cls(self.dim_z, self.step, nsliders, self.attrList).to(self.device)
        elif "NN" in walk_type:
            self.walk = WalkNonLinearW(self.dim_z, self.step, nsliders, self.attrList).to(self.device)
        else:
            raise NotImplementedError("Not implemented latent walk type: {}".format(walk_type))
        from latent2im_b200 import parallel
        parallel.broadcast_params(self.walk.parameters(), 0)   # ranks draw different numpy-global inits (transform_base.py:147)
        self.optimizers = torch.optim.Adam(self.walk.parameters(), lr=self.lr, betas=(0.5, 0.99))

    # ---- module construction ----------------------------------------------------------------
    def get_stylegan2_module(self):
        from .networks import Generator
        gen = Generator(constants.resolution, constants.DIM_Z, 8)
        if os.path.exists(constants.g_path):
            ckpt = torch.load(constants.g_path, map_location="cpu", weights_only=False)
            _report_load(gen.load_state_dict(ckpt["g_ema"], strict=False), constants.g_path)
        else:
            _missing_checkpoint("generator", constants.g_path)
        dtype = {"fp32": torch.float32, "bf16": torch.bfloat16}[getattr(constants, "compute_dtype", "bf16")]
        gen.set_native(dtype=dtype, max_batch=constants.BATCH_SIZE)
        module = _Module()
        module.netG = gen.to(self.device).eval()
        return module

    def get_reg_module(self):
        import torchvision
        model = torchvision.models.resnet50(weights=None)  # torch.hub needs network; same architecture
        model.fc = torch.nn.Linear(2048, 40)
        if os.path.exists(constants.reg_path):
            ckpt = torch.load(constants.reg_path, map_location="cpu", weights_only=False)
            model.load_state_dict(ckpt["model"])
        else:
            _missing_checkpoint("attribute regressor", constants.reg_path)
        model = model.to(self.device).eval()
        if getattr(constants, "reg_fold_bn", True):
            from latent2im_b200.regressor import fold_batchnorm
            model = fold_batchnorm(model, inplace=True)   # frozen eval-mode BN == constant affine map: folded into the convs
        if getattr(constants, "reg_amp", False):
            model = model.to(memory_format=torch.channels_last)
        for p in model.parameters():
            p.requires_grad_(False)  # frozen: only data gradients flow through R
        return model, None

    def get_attr_idx(self):
        return [self.attrTable[i] for i in self.attrList]

    # ---- hot path ----------------------------------------------------------------------------
    def get_w(self, z, is_single=False):
        w = self.module.netG.style(z)
        return [w] if is_single else [w] * (self.step + 1) * 2

    def get_logits(self, inputs_dict, reshape=True):
        latent = getattr(self.stylegan_opts, "latent", "w")
        if latent == "z":
            out, _ = self.module.netG([inputs_dict["z"]])
            return out
        w = inputs_dict["w"]
        w = torch.stack(list(w)).transpose(0, 1) if isinstance(w, (list, tuple)) else w.transpose(0, 1)
        out, _ = self.module.netG(w, input_is_latent=True)
        return out

    def get_w_new_tensor(self, multi_ws, alpha, layers=None, name=None, trainEmbed=False, index_=None):
        if isinstance(layers, str):  # train.py passes the raw --layers string (:174)
            layers = None if layers in ("", "None") else [int(i) for i in layers.split(",")]
        elif layers is not None:
            layers = [int(i) for i in layers]
        return self.walk(multi_ws, alpha=alpha, layers=layers)

    def _regress(self, logit):
        """The attribute regressor is a stock torchvision ResNet-50 (outside the accelerated path).  With
        ``constants.reg_amp`` it runs channels-last under bf16 autocast (about 2x faster at 1024 px); default = the
        reference's fp32 NCHW arithmetic."""
        if getattr(constants, "reg_amp", False):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return self.regressor(logit.contiguous(memory_format=torch.channels_last)).float()
        return self.regressor(logit)

    def get_reg_preds(self, logit):
        preds = self._regress(logit)[:, self.attrIdx]
        return preds.unsqueeze(1) if preds.ndim == 1 else preds

    def get_alphas(self, alpha_org, alpha_target):
        return alpha_target - alpha_org

    def get_bce_loss(self, pred, y, eps=1e-12):
        return -(y * pred.clamp(min=eps).log() + (1 - y) * (1 - pred).clamp(min=eps).log()).mean()

    def get_reg_loss(self, feed_dict):
        preds = self._regress(feed_dict["logit"])[:, self.attrIdx]
        return self.get_bce_loss(preds, feed_dict["alpha"].to(torch.double)).mean()

    # ---- optional loss terms (built only when requested) -----------------
    def get_discriminator(self):
        """The reference's ``StyleGAN(lr).netD`` (stylegan2.py:21-29): a Discriminator at the generator's resolution,
        never loaded from the checkpoint there; ``ckpt['d']`` is used here when the checkpoint file has one."""
        if self.module.netD is None:
            from .networks import Discriminator
            d = Discriminator(constants.resolution)
            if os.path.exists(constants.g_path):
                ckpt = torch.load(constants.g_path, map_location="cpu", weights_only=False)
                if "d" in ckpt:
                    _report_load(d.load_state_dict(ckpt["d"], strict=False), constants.g_path)
            d = d.to(self.device).eval()
            for p in d.parameters():
                p.requires_grad_(False)   # frozen: only the data gradient flows back to the walk
            self.module.netD = d
        return self.module.netD

    def get_vgg_module(self):
        """First eight modules of torchvision's VGG19 ``features`` = up to ``conv_4``, all the content loss reads
        (transform_base.py:427-453, 536-538).  ``constants.vgg_path`` (a ``vgg19`` state_dict) replaces the reference's
        ``pretrained=True`` download, which needs network."""
        if getattr(self, "vgg19", None) is None:
            import torchvision
            full = torchvision.models.vgg19(weights=None)
            path = getattr(constants, "vgg_path", "")
            if path and os.path.exists(path):
                full.load_state_dict(torch.load(path, map_location="cpu", weights_only=False))
            else:
                _missing_checkpoint("VGG19", path)
            # out-of-place ReLUs: the conv outputs are loss operands (the reference swaps them too, transform_base.py:439-441)
            vgg = torch.nn.Sequential(*[torch.nn.ReLU(inplace=False) if isinstance(m, torch.nn.ReLU) else m
                                        for m in full.features[:8]]).to(self.device).eval()
            for p in vgg.parameters():
                p.requires_grad_(False)
            self.vgg19 = vgg
            self._vgg_mean = torch.tensor([0.485, 0.456, 0.406], device=self.device).view(-1, 1, 1)
            self._vgg_std = torch.tensor([0.229, 0.224, 0.225], device=self.device).view(-1, 1, 1)
        return self.vgg19

    def get_content_loss(self, org_img, shifted_img):
        """MSE between the VGG19 activations of the original and the edited image after conv_1 .. conv_4
        (transform_base.py:427-453: the reference re-runs the growing prefix per layer; one pass gives the same values)."""
        vgg = self.get_vgg_module()
        a = (org_img.detach() - self._vgg_mean) / self._vgg_std
        b = (shifted_img - self._vgg_mean) / self._vgg_std
        losses = []
        for layer in vgg:
            a, b = layer(a), layer(b)
            if isinstance(layer, torch.nn.Conv2d):
                losses.append(torch.nn.functional.mse_loss(a.detach(), b))
        return losses

    def optimizeParametersAll(self, feed_dict, trainEmbed, updateGAN, no_content_loss=False, no_gan_loss=False):
        """transform_base.py:455-487: ``reg`` alone, or ``10 reg + 0.05 content + 0.05 gan`` for the terms not disabled."""
        self.optimizers.zero_grad()
        reg_loss = self.get_reg_loss(feed_dict)
        if no_content_loss and no_gan_loss:
            loss = reg_loss
        else:
            loss = 10 * reg_loss
            if not no_content_loss:
                content = self.get_content_loss(feed_dict["org"], feed_dict["logit"])
                loss = loss + 0.05 * (sum(content) / len(content))
            if not no_gan_loss:
                d_fake = self.get_discriminator()(feed_dict["logit"])
                loss = loss + 0.05 * torch.nn.functional.binary_cross_entropy_with_logits(d_fake, torch.ones_like(d_fake))
        loss.backward()
        # data parallel (one process per GPU): the only collective of the path, a no-op on a single GPU
        from latent2im_b200 import parallel
        parallel.allreduce_mean_grads(self.walk.parameters())
        self.optimizers.step()
        return loss

    def save_multi_models(self, save_path_w, save_path_gan, trainEmbed=False, updateGAN=False, single_transform_name=None):
        torch.save(self.walk, save_path_w + "_walk_module.ckpt")

    def load_multi_models(self, save_path_w, save_path_gan, trainEmbed=False, updateGAN=False, single_transform_name=None):
        import latent2im_b200
        latent2im_b200.install_dropin()  # reference pickles name graphs.stylegan_v2_real.transform_base.<Walk>
        self.walk = torch.load(save_path_w, map_location=self.device, weights_only=False)

    def clip_ims(self, ims):
        return np.uint8(np.clip(((ims + 1) / 2.0) * 255, 0, 255))

    def apply_alpha(self, graph_inputs, alpha_to_graph, layers=None, name=None, trainEmbed=False, index_=None,
                    given_w=None, return_uint8=False, cached_original=None):
        """``cached_original`` = (latent_w, out_zs, alpha_org) of an earlier call on the same z: the reference recomputes
        G(w) and R(G(w)) for every panel of a sweep (transform_base.py:554-603 inside the loop of :606-659); passing the
        first panel's result skips G-forward #1 and the regressor for the others."""
        with torch.no_grad():
            zs = graph_inputs["z"]
            if cached_original is not None:
                latent_w, out_zs, alpha_org = cached_original
            else:
                latent_w = given_w if given_w is not None else self.get_w(zs)
                out_zs = self.get_logits({"w": latent_w})
                alpha_org = self.get_reg_preds(out_zs)
            target = torch.as_tensor(np.asarray(alpha_to_graph), dtype=torch.float32, device=self.device)
            alpha_delta = self.get_alphas(alpha_org, target)
            if index_ is not None:
                col = index_ if len(self.attrIdx) == len(self.attrTable) else self.attrIdx.index(index_)
                alpha_delta[:, col] = target[:, 0] - alpha_org[:, col]
            latent_w_new = self.get_w_new_tensor(latent_w, alpha_delta, layers=layers, name=name, index_=index_)
            best_im_out = self.get_logits({"w": latent_w_new})
        self._last_original = (latent_w, out_zs, alpha_org)
        return best_im_out, alpha_org, out_zs

    def vis_multi_image_batch_alphas(self, graph_inputs, filename, alphas_to_graph, alphas_to_target, batch_start,
                                     layers=None, name=None, wgt=False, wmask=False, trainEmbed=False, computeL2=False,
                                     given_w=None, index_=None, cache_original=False):
        from latent2im_b200.utils import image
        zs_batch = graph_inputs["z"]
        panels = []
        cached = None
        for ag, _ in zip(alphas_to_graph, alphas_to_target):
            z = torch.Tensor(zs_batch).to(self.device)
            im, alpha_org, _ = self.apply_alpha({"z": z}, ag, name=name, layers=layers, given_w=given_w, index_=index_,
                                                cached_original=cached)
            if cache_original:
                cached = self._last_original
            u8 = torch.empty(im.shape[0], im.shape[2], im.shape[3], 3, device=im.device, dtype=torch.uint8)
            nt.check(nt.load().l2i_image_to_uint8(u8.data_ptr(), im.contiguous().data_ptr(), im.shape[0], im.shape[2],
                                                  im.shape[3], nt.stream_ptr(im.device)), "image_to_uint8")
            panels.append(u8.cpu().numpy())
        for ii in range(zs_batch.shape[0]):
            a = alpha_org[ii, index_].item() if (index_ is not None and len(self.attrList) > 1) else alpha_org[ii].flatten()[0].item()
            ims = np.stack([p[ii] for p in panels], axis=0)
            fn = filename + "_sample{}".format(ii + batch_start) + ("_wgt" if wgt else "") + "_%.2f" % a
            image.save_im(image.imgrid(ims, cols=len(alphas_to_graph)), fn)


class PixelTransform(TransformGraph):
    def __init__(self, *args, **kwargs):
        TransformGraph.__init__(self, *args, **kwargs)
