"""CPU: the oracle's two op restatements against closed-form cases and against each other."""
import itertools

import pytest
import torch

from oracle.ops import (fused_bias_act_ref, fused_leaky_relu_bwd_ref, fused_leaky_relu_ref, make_fir_kernel,
                        upfirdn2d_nhwc_ref, upfirdn2d_out_size, upfirdn2d_pixel_ref, upfirdn2d_ref)


def test_identity_kernel_is_identity():
    x = torch.randn(2, 3, 7, 5, dtype=torch.float64)
    k = torch.ones(1, 1, dtype=torch.float64)
    assert torch.equal(upfirdn2d_ref(x, k), x)


def test_zero_insert_upsample():
    x = torch.arange(12, dtype=torch.float64).reshape(1, 1, 3, 4)
    y = upfirdn2d_ref(x, torch.ones(1, 1, dtype=torch.float64), up=2)
    assert y.shape == (1, 1, 6, 8)
    assert torch.equal(y[:, :, ::2, ::2], x)
    assert y[:, :, 1::2].abs().sum() == 0 and y[:, :, :, 1::2].abs().sum() == 0


def test_blur_preserves_constant_interior():
    k = make_fir_kernel([1, 3, 3, 1], dtype=torch.float64)
    x = torch.full((1, 1, 12, 12), 2.5, dtype=torch.float64)
    y = upfirdn2d_ref(x, k, pad=(2, 1))
    assert y.shape == (1, 1, 12, 12)
    assert torch.allclose(y[:, :, 3:-3, 3:-3], torch.full_like(y[:, :, 3:-3, 3:-3], 2.5))


def test_upsample_kernel_gain_keeps_mean():
    k = make_fir_kernel([1, 3, 3, 1], gain=4, dtype=torch.float64)
    x = torch.full((1, 1, 8, 8), 1.0, dtype=torch.float64)
    y = upfirdn2d_ref(x, k, up=2, pad=(2, 1))
    assert y.shape == (1, 1, 16, 16)
    assert torch.allclose(y[:, :, 4:-4, 4:-4], torch.ones_like(y[:, :, 4:-4, 4:-4]))


def test_negative_pad_crops():
    x = torch.randn(1, 1, 6, 6, dtype=torch.float64)
    y = upfirdn2d_ref(x, torch.ones(1, 1, dtype=torch.float64), pad=(-1, -2))
    assert torch.equal(y, x[:, :, 1:-2, 1:-2])


CONFIGS = [
    # (in_h, in_w, kh, kw, up, down, pad0, pad1)
    (9, 9, 4, 4, 1, 1, 1, 1), (8, 8, 4, 4, 2, 1, 2, 1), (16, 16, 4, 4, 1, 2, 1, 1), (7, 5, 4, 4, 1, 1, 2, 2),
    (5, 6, 3, 3, 1, 1, 1, 1), (4, 4, 2, 2, 2, 1, 0, 0), (6, 6, 2, 2, 1, 2, 0, 0), (5, 7, 4, 3, 2, 2, 1, 2),
    (3, 3, 4, 4, 3, 2, 2, 3), (6, 4, 1, 1, 1, 1, -1, -1), (1, 1, 4, 4, 2, 1, 2, 1),
]


@pytest.mark.parametrize("cfg", CONFIGS)
@pytest.mark.parametrize("minor", [1, 3])
def test_slice_version_matches_cuda_index_maths(cfg, minor):
    in_h, in_w, kh, kw, up, down, p0, p1 = cfg
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, in_h, in_w, minor, generator=g, dtype=torch.float64)
    k = torch.randn(kh, kw, generator=g, dtype=torch.float64)
    a = upfirdn2d_nhwc_ref(x, k, up, up, down, down, p0, p1, p0, p1)
    b = upfirdn2d_pixel_ref(x, k, up, up, down, down, p0, p1, p0, p1)
    oh, ow = upfirdn2d_out_size(in_h, in_w, kh, kw, up, up, down, down, p0, p1, p0, p1)
    assert a.shape == (2, max(oh, 0), max(ow, 0), minor) == b.shape
    assert torch.allclose(a, b, atol=1e-12)


def test_upfirdn_backward_is_transposed_op():
    """The reference's backward = same op with up/down swapped and the flipped kernel
    (op/upfirdn2d.py:30-41, 110-113); check that against autograd of the oracle."""
    for in_h, in_w, kh, kw, up, down, p0, p1 in CONFIGS[:5]:
        x = torch.randn(1, 2, in_h, in_w, dtype=torch.float64, requires_grad=True)
        k = torch.randn(kh, kw, dtype=torch.float64)
        y = upfirdn2d_ref(x, k, up=up, down=down, pad=(p0, p1))
        gy = torch.randn_like(y)
        (gx,) = torch.autograd.grad(y, x, gy)
        out_h, out_w = y.shape[2:]
        gp = (kw - p0 - 1, in_w * up - out_w * down + p0 - up + 1, kh - p0 - 1, in_h * up - out_h * down + p0 - up + 1)
        g4 = upfirdn2d_nhwc_ref(gy.reshape(-1, out_h, out_w, 1), torch.flip(k, [0, 1]), down, down, up, up, *gp)
        assert torch.allclose(g4.reshape(gx.shape), gx, atol=1e-12)


def test_fused_bias_act_cases():
    x = torch.tensor([[-2.0, 3.0], [0.5, -0.25]], dtype=torch.float64)
    b = torch.tensor([1.0, -1.0], dtype=torch.float64)
    y = fused_leaky_relu_ref(x, b)
    v = x + b
    exp = torch.where(v > 0, v, 0.2 * v) * 2 ** 0.5
    assert torch.allclose(y, exp)
    assert torch.equal(fused_bias_act_ref(x, None, None, 1, 0, 0.2, 1.0), x)
    assert fused_bias_act_ref(x, None, None, 3, 2, 0.2, 1.0).abs().sum() == 0
    ref = torch.tensor([[1.0, -1.0], [-1.0, 1.0]], dtype=torch.float64)
    g = fused_bias_act_ref(x, None, ref, 3, 1, 0.2, 2.0)
    assert torch.allclose(g, torch.where(ref > 0, x, 0.2 * x) * 2.0)


def test_fused_lrelu_backward_matches_autograd():
    x = torch.randn(3, 5, 4, 4, dtype=torch.float64, requires_grad=True)
    b = torch.randn(5, dtype=torch.float64, requires_grad=True)
    y = fused_leaky_relu_ref(x, b)
    gy = torch.randn_like(y)
    gx, gb = torch.autograd.grad(y, (x, b), gy)
    gi, gbias = fused_leaky_relu_bwd_ref(gy, y.detach())
    assert torch.allclose(gi, gx) and torch.allclose(gbias, gb)
