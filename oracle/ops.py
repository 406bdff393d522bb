"""Oracle restatement of the reference's two native ops (test infrastructure only).

* ``upfirdn2d_ref``      follows reference op/upfirdn2d.py:152-186 (``upfirdn2d_native``:
  zero-insert upsample -> pad / negative-pad crop -> FIR with the flipped kernel ->
  decimate) and is cross-checked in tests against the per-output-pixel index
  arithmetic of op/upfirdn2d_kernel.cu:85-129,167-168 (``upfirdn2d_pixel_ref``).
* ``fused_bias_act_ref`` follows op/fused_bias_act_kernel.cu:25-47 (bias broadcast on
  dim 1, ``act*10+grad`` switch, final ``* scale``) and op/fused_act.py:28-39 for the
  bias-gradient reduction.

Everything works in the dtype of its input (float64 for the arbiter runs).
"""
from __future__ import annotations

import torch


def make_fir_kernel(taps, gain: float = 1.0, dtype=torch.float32) -> torch.Tensor:
    """Reference networks.py:19-27 ``make_kernel``: outer product of a 1-D tap list,
    normalised to unit sum (then optionally multiplied by ``gain`` = factor**2)."""
    k = torch.tensor(taps, dtype=torch.float32)
    if k.ndim == 1:
        k = k[None, :] * k[:, None]
    k = k / k.sum()
    return (k * gain).to(dtype)


def upfirdn2d_ref(x: torch.Tensor, kernel: torch.Tensor, up=1, down=1, pad=(0, 0)) -> torch.Tensor:
    """NCHW front end, reference op/upfirdn2d.py:144-149 (pad = (pad0, pad1) used for x and y)."""
    b, c, h, w = x.shape
    out = upfirdn2d_nhwc_ref(
        x.reshape(b * c, h, w, 1), kernel, up, up, down, down, pad[0], pad[1], pad[0], pad[1]
    )
    return out.reshape(b, c, out.shape[1], out.shape[2])


def upfirdn2d_nhwc_ref(x, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
    """``[major, in_h, in_w, minor]`` layout of the native op (op/upfirdn2d.cpp:12-22).

    Steps follow op/upfirdn2d.py:152-186.  The FIR is evaluated as a sum of shifted
    slices (a correlation with the flipped kernel, i.e. a true convolution)."""
    major, in_h, in_w, minor = x.shape
    kh, kw = kernel.shape
    kernel = kernel.to(x.dtype)

    # zero-insert upsample: sample (y, x) lands on (y*up_y, x*up_x); zeros are appended after it
    up = x.new_zeros(major, in_h * up_y, in_w * up_x, minor)
    up[:, ::up_y, ::up_x, :] = x

    # positive pads add zeros, negative pads crop
    ph, pw = in_h * up_y + pad_y0 + pad_y1, in_w * up_x + pad_x0 + pad_x1
    padded = x.new_zeros(major, max(ph, 0), max(pw, 0), minor)
    sy0, sx0 = max(-pad_y0, 0), max(-pad_x0, 0)
    sy1, sx1 = in_h * up_y - max(-pad_y1, 0), in_w * up_x - max(-pad_x1, 0)
    dy0, dx0 = max(pad_y0, 0), max(pad_x0, 0)
    if sy1 > sy0 and sx1 > sx0:
        padded[:, dy0:dy0 + (sy1 - sy0), dx0:dx0 + (sx1 - sx0), :] = up[:, sy0:sy1, sx0:sx1, :]

    oh, ow = ph - kh + 1, pw - kw + 1
    flipped = torch.flip(kernel, [0, 1])
    acc = x.new_zeros(major, max(oh, 0), max(ow, 0), minor)
    if oh > 0 and ow > 0:
        for i in range(kh):
            for j in range(kw):
                acc = acc + flipped[i, j] * padded[:, i:i + oh, j:j + ow, :]
    return acc[:, ::down_y, ::down_x, :]


def upfirdn2d_out_size(in_h, in_w, kh, kw, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
    """op/upfirdn2d_kernel.cu:167-168."""
    out_h = (in_h * up_y + pad_y0 + pad_y1 - kh + down_y) // down_y
    out_w = (in_w * up_x + pad_x0 + pad_x1 - kw + down_x) // down_x
    return out_h, out_w


def _floor_div(a: int, b: int) -> int:
    return a // b  # Python's // already floors; op/upfirdn2d_kernel.cu:18-26 emulates it in C


def upfirdn2d_pixel_ref(x, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
    """Scalar-loop restatement of the CUDA kernel's per-output-pixel arithmetic
    (op/upfirdn2d_kernel.cu:71-81 tap flip, :85-88 and :114-129 gather).  Small inputs only."""
    major, in_h, in_w, minor = x.shape
    kh, kw = kernel.shape
    out_h, out_w = upfirdn2d_out_size(in_h, in_w, kh, kw, up_x, up_y, down_x, down_y,
                                      pad_x0, pad_x1, pad_y0, pad_y1)
    out = x.new_zeros(major, out_h, out_w, minor)
    k = kernel.to(x.dtype)
    for oy in range(out_h):
        mid_y = oy * down_y + up_y - 1 - pad_y0
        in_y0 = _floor_div(mid_y, up_y)
        ky0 = (in_y0 + 1) * up_y - mid_y - 1
        for ox in range(out_w):
            mid_x = ox * down_x + up_x - 1 - pad_x0
            in_x0 = _floor_div(mid_x, up_x)
            kx0 = (in_x0 + 1) * up_x - mid_x - 1
            v = x.new_zeros(major, minor)
            # the CUDA kernel walks kernel_h/up_y x kernel_w/up_x taps of the *flipped* kernel;
            # taps beyond the true kernel extent are zero (kernel is zero-padded to the template size)
            ty = 0
            while ky0 + ty * up_y < kh:
                tx = 0
                while kx0 + tx * up_x < kw:
                    iy, ix = in_y0 + ty, in_x0 + tx
                    if 0 <= iy < in_h and 0 <= ix < in_w:
                        fy, fx = ky0 + ty * up_y, kx0 + tx * up_x
                        v = v + x[:, iy, ix, :] * k[kh - 1 - fy, kw - 1 - fx]
                    tx += 1
                ty += 1
            out[:, oy, ox, :] = v
    return out


def fused_bias_act_ref(x, bias=None, ref=None, act=3, grad=0, alpha=0.2, scale=2 ** 0.5):
    """op/fused_bias_act_kernel.cu:25-47.  ``bias`` is indexed by dim 1 of ``x``."""
    v = x
    if bias is not None and bias.numel():
        shape = [1, -1] + [1] * (x.ndim - 2)
        v = v + bias.to(x.dtype).reshape(shape)
    code = act * 10 + grad
    if code in (10, 11):
        y = v
    elif code in (12, 32):
        y = torch.zeros_like(v)
    elif code == 30:
        y = torch.where(v > 0, v, v * alpha)
    elif code == 31:
        y = torch.where(ref > 0, v, v * alpha)
    else:  # the kernel's ``default:`` label falls into case 10
        y = v
    return y * scale


def fused_leaky_relu_ref(x, bias, negative_slope=0.2, scale=2 ** 0.5):
    """op/fused_act.py:51-59, 85-86 forward."""
    return fused_bias_act_ref(x, bias, None, 3, 0, negative_slope, scale)


def fused_leaky_relu_bwd_ref(grad_out, out, negative_slope=0.2, scale=2 ** 0.5):
    """op/fused_act.py:19-39: grad_input via (act=3, grad=1, ref=saved output); grad_bias = sum over
    all dims but 1."""
    gi = fused_bias_act_ref(grad_out, None, out, 3, 1, negative_slope, scale)
    dims = [0] + list(range(2, gi.ndim))
    return gi, gi.sum(dims)
