"""``upfirdn2d`` over the sm_100a FIR resampling kernel.

Same call signature and autograd behaviour as the reference's op/upfirdn2d.py:144-149 (forward,
backward = the same op with up/down swapped and the flipped kernel, double backward), executed by
``l2i_upfirdn2d`` from the C-ABI library.
"""
import torch
from torch.autograd import Function

from latent2im_b200 import _native as nt


def _native_upfirdn2d(x4, kernel, up, down, pad):
    """x4: [major, h, w, minor] contiguous.  Returns [major, out_h, out_w, minor]."""
    nt.require_cuda(x4, "input")
    nt.require_cuda(kernel, "kernel")
    major, in_h, in_w, minor = x4.shape
    kh, kw = kernel.shape
    (up_x, up_y), (down_x, down_y), (px0, px1, py0, py1) = up, down, pad
    out_h = max((in_h * up_y + py0 + py1 - kh + down_y) // down_y, 0)
    out_w = max((in_w * up_x + px0 + px1 - kw + down_x) // down_x, 0)
    y = torch.empty(major, out_h, out_w, minor, device=x4.device, dtype=x4.dtype)
    k32 = kernel.to(torch.float32).contiguous()
    with torch.cuda.device(x4.device):
        rc = nt.load().l2i_upfirdn2d(nt.ptr(y), nt.ptr(x4), nt.ptr(k32), major, in_h, in_w, minor, kh, kw,
                                     up_x, up_y, down_x, down_y, px0, px1, py0, py1,
                                     nt.dtype_code(x4.dtype), nt.stream_ptr(x4.device))
    nt.check(rc, "upfirdn2d")
    return y


class _UpFirDnGrad(Function):
    @staticmethod
    def forward(ctx, grad_out, kernel, kernel_flipped, up, down, pad, g_pad, in_shape, out_hw):
        g4 = grad_out.reshape(-1, out_hw[0], out_hw[1], 1).contiguous()
        gi = _native_upfirdn2d(g4, kernel_flipped, down, up, g_pad)
        ctx.save_for_backward(kernel)
        ctx.cfg = (up, down, pad, in_shape, out_hw)
        return gi.view(in_shape)

    @staticmethod
    def backward(ctx, gg_in):
        kernel, = ctx.saved_tensors
        up, down, pad, in_shape, out_hw = ctx.cfg
        gg4 = gg_in.reshape(-1, in_shape[2], in_shape[3], 1).contiguous()
        gg_out = _native_upfirdn2d(gg4, kernel, up, down, pad)
        return (gg_out.view(in_shape[0], in_shape[1], out_hw[0], out_hw[1]),) + (None,) * 8


class _UpFirDn(Function):
    @staticmethod
    def forward(ctx, x, kernel, up, down, pad):
        b, c, in_h, in_w = x.shape
        kh, kw = kernel.shape
        (up_x, up_y), (down_x, down_y), (px0, px1, py0, py1) = up, down, pad
        out = _native_upfirdn2d(x.reshape(-1, in_h, in_w, 1).contiguous(), kernel, up, down, pad)
        out_h, out_w = out.shape[1], out.shape[2]
        # padding of the transposed op (op/upfirdn2d.py:110-113)
        g_pad = (kw - px0 - 1, in_w * up_x - out_w * down_x + px0 - up_x + 1,
                 kh - py0 - 1, in_h * up_y - out_h * down_y + py0 - up_y + 1)
        ctx.save_for_backward(kernel, torch.flip(kernel, [0, 1]))
        ctx.cfg = (up, down, pad, g_pad, tuple(x.shape), (out_h, out_w))
        return out.view(-1, c, out_h, out_w)

    @staticmethod
    def backward(ctx, grad_out):
        kernel, kernel_flipped = ctx.saved_tensors
        up, down, pad, g_pad, in_shape, out_hw = ctx.cfg
        gi = _UpFirDnGrad.apply(grad_out, kernel, kernel_flipped, up, down, pad, g_pad, in_shape, out_hw)
        return gi, None, None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    return _UpFirDn.apply(input, kernel, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]))
