"""``fused_leaky_relu`` / ``FusedLeakyReLU`` over the sm_100a bias-act kernels.

Mirrors the interface of the reference's op/fused_act.py:51-86 (same names, argument meaning, the
``bias`` parameter name used by checkpoints) but the work is done by ``l2i_fused_bias_act`` and
``l2i_fused_leaky_relu_bwd`` from the C-ABI library; CPU tensors raise like the reference's
``CHECK_CUDA`` (op/fused_bias_act.cpp:13-14).
"""
import torch
from torch import nn
from torch.autograd import Function

from latent2im_b200 import _native as nt


def _bias_geometry(x):
    """bias is indexed by dim 1 (op/fused_bias_act_kernel.cu:65-70): step = prod(shape[2:])."""
    step = 1
    for s in x.shape[2:]:
        step *= s
    return step, (x.shape[1] if x.ndim > 1 else 1)


def _bias_act(x, bias, ref, act, grad, alpha, scale):
    nt.require_cuda(x, "input")
    if bias is not None:
        nt.require_cuda(bias, "bias")
    x = x.contiguous()
    y = torch.empty_like(x)
    step, size_b = _bias_geometry(x)
    b = None
    if bias is not None and bias.numel():
        b = bias.to(x.dtype).contiguous()
        size_b = b.numel()
    r = ref.contiguous() if ref is not None and ref.numel() else None
    with torch.cuda.device(x.device):
        rc = nt.load().l2i_fused_bias_act(nt.ptr(y), nt.ptr(x), nt.ptr(b), nt.ptr(r), x.numel(), step,
                                          size_b if b is not None else 0, act, grad, float(alpha),
                                          float(scale), nt.dtype_code(x.dtype), nt.stream_ptr(x.device))
    nt.check(rc, "fused_bias_act")
    return y


class _LeakyBiasActGrad(Function):
    """grad_input, grad_bias of the forward; itself differentiable (double backward)."""

    @staticmethod
    def forward(ctx, grad_out, out, alpha, scale):
        ctx.save_for_backward(out)
        ctx.alpha, ctx.scale = alpha, scale
        g = grad_out.contiguous()
        gi = torch.empty_like(g)
        step, size_b = _bias_geometry(g)
        outer = g.shape[0] if g.ndim > 1 else 1
        gb = torch.empty(size_b, device=g.device, dtype=torch.float32)
        with torch.cuda.device(g.device):
            rc = nt.load().l2i_fused_leaky_relu_bwd(nt.ptr(gi), nt.ptr(gb), nt.ptr(g), nt.ptr(out.contiguous()),
                                                    outer, size_b, step, float(alpha), float(scale),
                                                    nt.dtype_code(g.dtype), nt.stream_ptr(g.device))
        nt.check(rc, "fused_leaky_relu_bwd")
        return gi, gb.to(g.dtype)

    @staticmethod
    def backward(ctx, gg_in, gg_bias):
        out, = ctx.saved_tensors
        gg_out = _bias_act(gg_in, gg_bias, out, 3, 1, ctx.alpha, ctx.scale)
        return gg_out, None, None, None


class _LeakyBiasAct(Function):
    @staticmethod
    def forward(ctx, x, bias, alpha, scale):
        out = _bias_act(x, bias, None, 3, 0, alpha, scale)
        ctx.save_for_backward(out)
        ctx.alpha, ctx.scale = alpha, scale
        return out

    @staticmethod
    def backward(ctx, grad_out):
        out, = ctx.saved_tensors
        gi, gb = _LeakyBiasActGrad.apply(grad_out, out, ctx.alpha, ctx.scale)
        return gi, gb, None, None


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
    return _LeakyBiasAct.apply(input, bias, negative_slope, scale)


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)
