"""GPU: the native-op drop-ins (through the C ABI) against the CPU oracle."""
import pytest
import torch

from oracle.ops import fused_leaky_relu_bwd_ref, fused_leaky_relu_ref, upfirdn2d_ref

pytestmark = pytest.mark.gpu

UPFIRDN_CASES = [
    # shape, kernel hw, up, down, pad
    ((2, 3, 9, 9), (4, 4), 1, 1, (1, 1)), ((2, 3, 8, 8), (4, 4), 2, 1, (2, 1)), ((2, 3, 16, 16), (4, 4), 1, 2, (1, 1)),
    ((1, 2, 7, 5), (4, 4), 1, 1, (2, 2)), ((1, 2, 33, 70), (4, 4), 2, 1, (2, 1)), ((1, 1, 12, 12), (3, 3), 1, 1, (1, 1)),
    ((1, 2, 40, 130), (4, 4), 1, 1, (1, 1)), ((1, 1, 5, 7), (4, 3), 2, 2, (1, 2)), ((1, 1, 3, 3), (4, 4), 3, 2, (2, 3)),
    ((1, 1, 6, 4), (1, 1), 1, 1, (-1, -1)), ((2, 1, 1, 1), (4, 4), 2, 1, (2, 1)), ((1, 1, 8, 8), (8, 8), 1, 4, (3, 4)),
]


@pytest.mark.parametrize("case", UPFIRDN_CASES)
def test_upfirdn2d_forward_backward(case):
    from latent2im_b200.graphs.stylegan_v2_real.op import upfirdn2d
    shape, khw, up, down, pad = case
    g = torch.Generator().manual_seed(0)
    x = torch.randn(shape, generator=g, dtype=torch.float64)
    k = torch.randn(khw, generator=g, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    ref = upfirdn2d_ref(xr, k, up=up, down=down, pad=pad)
    gy = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    (gref,) = torch.autograd.grad(ref, xr, gy)
    xc = x.float().cuda().requires_grad_(True)
    out = upfirdn2d(xc, k.float().cuda(), up=up, down=down, pad=pad)
    assert out.shape == ref.shape
    assert torch.allclose(out.detach().cpu().double(), ref.detach(), atol=2e-5, rtol=1e-5)
    (gx,) = torch.autograd.grad(out, xc, gy.float().cuda())
    assert torch.allclose(gx.cpu().double(), gref, atol=2e-5, rtol=1e-5)


def test_upfirdn2d_nhwc_minor_and_empty():
    from latent2im_b200.graphs.stylegan_v2_real.op.upfirdn2d import _native_upfirdn2d
    from oracle.ops import upfirdn2d_nhwc_ref
    x = torch.randn(2, 6, 5, 3, dtype=torch.float64)
    k = torch.randn(4, 4, dtype=torch.float64)
    ref = upfirdn2d_nhwc_ref(x, k, 2, 2, 1, 1, 2, 1, 2, 1)
    out = _native_upfirdn2d(x.float().cuda(), k.float().cuda(), (2, 2), (1, 1), (2, 1, 2, 1))
    assert torch.allclose(out.cpu().double(), ref, atol=2e-5)
    empty = _native_upfirdn2d(torch.zeros(0, 4, 4, 1).cuda(), k.float().cuda(), (1, 1), (1, 1), (1, 1, 1, 1))
    assert empty.shape == (0, 3, 3, 1)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(2, 8, 5, 5), (3, 16), (1, 4, 33, 17), (2, 6, 16, 16), (4, 7)])
def test_fused_leaky_relu_forward_backward(dtype, shape):
    from latent2im_b200.graphs.stylegan_v2_real.op import fused_leaky_relu
    g = torch.Generator().manual_seed(1)
    x = torch.randn(shape, generator=g).to(dtype)
    b = torch.randn(shape[1], generator=g).to(dtype)
    ref = fused_leaky_relu_ref(x.double(), b.double())
    xc, bc = x.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    out = fused_leaky_relu(xc, bc)
    tol = 1e-6 if dtype == torch.float32 else 2e-2
    assert torch.allclose(out.detach().cpu().double(), ref, atol=tol, rtol=tol)
    gy = torch.randn(shape, generator=g).to(dtype)
    gx, gb = torch.autograd.grad(out, (xc, bc), gy.cuda())
    rgx, rgb = fused_leaky_relu_bwd_ref(gy.double(), out.detach().cpu().double())
    assert torch.allclose(gx.cpu().double(), rgx, atol=tol, rtol=tol)
    assert torch.allclose(gb.cpu().double(), rgb, atol=tol * 50, rtol=tol * 5)


def test_fused_leaky_relu_module_and_double_backward():
    from latent2im_b200.graphs.stylegan_v2_real.op import FusedLeakyReLU
    m = FusedLeakyReLU(4).cuda()
    x = torch.randn(2, 4, 3, 3, device="cuda", requires_grad=True)
    y = m(x)
    (gx,) = torch.autograd.grad(y.sum(), x, create_graph=True)
    gx.sum().backward()  # runs the double-backward kernel path
    assert x.grad is not None and x.grad.abs().sum() == 0  # piecewise linear: second derivative is zero


def test_large_index_bias_act():
    """> 2^31 elements: the reference's int32 size_x overflows here."""
    from latent2im_b200.graphs.stylegan_v2_real.op import fused_leaky_relu
    n = 2 ** 31 + 4096
    x = torch.zeros(1, 2, n // 2, device="cuda", dtype=torch.bfloat16)
    x[0, 1, -1] = -5.0
    b = torch.tensor([1.0, 2.0], device="cuda", dtype=torch.bfloat16)
    y = fused_leaky_relu(x, b)
    assert abs(y[0, 0, 0].item() - 2 ** 0.5) < 1e-2
    assert abs(y[0, 1, -1].item() - (-3.0 * 0.2 * 2 ** 0.5)) < 1e-2
    assert abs(y[0, 1, 0].item() - 2 * 2 ** 0.5) < 2e-2


def test_latent_kernels():
    import ctypes
    from latent2im_b200 import _native as nt
    from oracle.walks import walk_linear_ref
    lib = nt.load()
    g = torch.Generator().manual_seed(0)
    B, L, D, A = 5, 6, 64, 3
    ws = torch.randn(B, L, D, generator=g)
    alpha = torch.randn(B, A, generator=g)
    w = torch.randn(A, L, D, generator=g)
    ref = torch.stack(walk_linear_ref([ws[:, i].double() for i in range(L)], alpha.double(), w.double(), layers=[0, 2, 5]), 1)
    out = torch.empty(B, L, D, device="cuda")
    wsc, alc, wc = ws.cuda(), alpha.cuda(), w.cuda()  # keep the device buffers alive across the call
    mask = (1 << 0) | (1 << 2) | (1 << 5)
    nt.check(lib.l2i_walk_linear_fwd(out.data_ptr(), wsc.data_ptr(), L * D, D, alc.data_ptr(), wc.data_ptr(),
                                     B, A, L, D, mask, nt.stream_ptr()), "walk")
    assert torch.allclose(out.cpu().double(), ref, atol=1e-5)
    # linear + pixel norm
    x = torch.randn(B, D, generator=g)
    W = torch.randn(48, D, generator=g)
    bias = torch.randn(48, generator=g)
    y = torch.empty(B, 48, device="cuda")
    xc, Wc, bc = x.cuda(), W.cuda(), bias.cuda()
    nt.check(lib.l2i_linear_fwd(y.data_ptr(), 48, xc.data_ptr(), D, Wc.data_ptr(), bc.data_ptr(), B, 48, D,
                                0.5, 2.0, 1, 0.2, 1.5, nt.stream_ptr()), "linear")
    v = (x.double() @ W.double().t()) * 0.5 + bias.double() * 2.0
    v = torch.where(v > 0, v, 0.2 * v) * 1.5
    assert torch.allclose(y.cpu().double(), v, atol=1e-4)
