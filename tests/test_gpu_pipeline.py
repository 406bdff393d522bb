"""GPU: the public end-to-end call (host z / alpha in, host uint8 panels out) that bench.py's ``e2e`` number times,
against the CPU oracle: mapping -> linear walk -> synthesis -> clip((x+1)/2*255) truncated (transform_base.py:554-626)."""
import numpy as np
import pytest
import torch

from latent2im_b200.synthetic import load_synthetic, synthetic_noise, synthetic_walk_w, synthetic_z
from oracle import GeneratorSpec, generator_forward_ref, mapping_ref
from oracle.generator import clip_to_uint8_ref
from oracle.walks import walk_linear_ref

pytestmark = pytest.mark.gpu


def _setup(size, dim, n_mlp, batch, dtype):
    from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
    from latent2im_b200.graphs.stylegan_v2_real.transform_base import WalkLinearMultiW
    from latent2im_b200.pipeline import EditPipeline
    spec = GeneratorSpec(size=size, style_dim=dim, n_mlp=n_mlp)
    gen = load_synthetic(Generator(size, dim, n_mlp), seed=0)
    sd = {k: v.double() for k, v in gen.state_dict().items()}
    gen = gen.cuda().eval()
    gen.set_native(dtype=dtype, max_batch=batch)
    np.random.seed(0)
    walk = WalkLinearMultiW(dim, spec.log_size - 2, 1, ["Smiling"]).cuda()
    w0 = synthetic_walk_w(1, spec.n_latent, dim, seed=0)
    with torch.no_grad():
        walk.w.copy_(w0.cuda())
    return spec, sd, w0, EditPipeline(gen, walk, batch, n_attr=1, device=torch.device("cuda"))


def _oracle_u8(spec, sd, w0, z, alpha, noise):
    w = mapping_ref(sd, torch.tensor(z).double(), spec)
    lat = torch.stack(walk_linear_ref([w] * spec.n_latent, alpha.double(), w0.double()), 1)
    img = generator_forward_ref(sd, lat, noise, spec)
    return img, clip_to_uint8_ref(img.float()).permute(0, 2, 3, 1).numpy()


@pytest.mark.parametrize("batch", [1, 3])
def test_edit_fp32_matches_oracle_uint8(batch):
    spec, sd, w0, pipe = _setup(32, 64, 2, batch, torch.float32)
    noise = synthetic_noise(spec.num_layers, batch)
    z = synthetic_z(batch, 5, 64)                       # float64 host array, as graph_util.z_sample returns it
    alpha = torch.linspace(-0.5, 1.0, batch).reshape(batch, 1)
    out = pipe.edit(z, alpha, noise=[n.cuda() for n in noise]).copy()
    img, ref = _oracle_u8(spec, sd, w0, z.astype(np.float32), alpha, noise)
    assert out.shape == (batch, 32, 32, 3) and out.dtype == np.uint8
    # truncating cast: values within 1e-3 of an integer boundary may land on either side
    frac = ((img.float().permute(0, 2, 3, 1).numpy() + 1) / 2 * 255) % 1.0
    safe = (frac > 2e-1) & (frac < 1 - 2e-1)
    assert np.abs(out.astype(int) - ref.astype(int)).max() <= 1
    assert np.array_equal(out[safe], ref[safe])


def test_async_double_buffering_returns_every_result():
    """sync=False: the device->host copy of call i overlaps call i+1; results of consecutive calls must not be mixed up."""
    spec, sd, w0, pipe = _setup(16, 32, 1, 2, torch.float32)
    noise = [n.cuda() for n in synthetic_noise(spec.num_layers, 2)]
    zs = [synthetic_z(2, s, 32) for s in range(4)]
    alpha = torch.full((2, 1), 0.3)
    expect = [pipe.edit(z, alpha, noise=noise, sync=True).copy() for z in zs]
    got = []
    for z in zs:
        view = pipe.edit(z, alpha, noise=noise, sync=False)
        pipe.wait()
        got.append(view.copy())
    for a, b in zip(expect, got):
        assert np.array_equal(a, b)
    assert not np.array_equal(expect[0], expect[1])
    assert pipe.h2d_bytes == 2 * 32 * 4 + 2 * 4 and pipe.d2h_bytes == 2 * 16 * 16 * 3


def test_async_calls_in_flight_keep_their_own_inputs():
    """Several sync=False calls with DIFFERENT z / alpha are issued back to back before any wait: every result must be the
    image of its own inputs (the pinned input buffers are double-buffered and guarded by an event; a single pinned
    buffer would let call i synthesize from call i+1's latents)."""
    spec, sd, w0, pipe = _setup(64, 64, 1, 4, torch.float32)
    noise = [n.cuda() for n in synthetic_noise(spec.num_layers, 4)]
    zs = [synthetic_z(4, 10 + s, 64) for s in range(6)]
    alphas = [torch.full((4, 1), 0.1 * s) for s in range(6)]
    expect = [pipe.edit(z, a, noise=noise, sync=True).copy() for z, a in zip(zs, alphas)]
    # keep the GPU busy so that the host really runs ahead of the queued host->device copies
    big = torch.randn(8192, 8192, device="cuda")
    for _ in range(8):
        big = big @ big.t() * 1e-4
    got = []
    for i, (z, a) in enumerate(zip(zs, alphas)):
        view = pipe.edit(z, a, noise=noise, sync=False)
        if i >= 1:                                     # result i-1 lives in the other output buffer: read it one call late
            pipe.copy_done[(i - 1) & 1].synchronize()
            got.append(prev.copy())
        prev = view
    pipe.wait()
    got.append(prev.copy())
    for i, (a, b) in enumerate(zip(expect, got)):
        assert np.array_equal(a, b), i
    assert not np.array_equal(expect[0], expect[1])


def test_pipeline_needs_cuda():
    from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
    from latent2im_b200.pipeline import EditPipeline
    with pytest.raises(RuntimeError):
        EditPipeline(Generator(16, 32, 1), None, 2, device=torch.device("cpu"))


def test_noise_drawn_ahead_consumes_the_reference_stream():
    """EditPipeline draws call i+1's noise on a side stream during call i; a seeded run must still see exactly the
    tensors ``NoiseInjection`` would draw (same generator, order, shapes) - checked against explicit-noise runs."""
    spec, sd, w0, pipe = _setup(16, 32, 1, 2, torch.float32)
    z = torch.tensor(synthetic_z(2, 1, 32), dtype=torch.float32).cuda()
    alpha = torch.full((2, 1), 0.3, device="cuda")
    torch.manual_seed(77)
    pipe._next_noise = None
    a1 = pipe.edit_device(z, alpha).clone()
    a2 = pipe.edit_device(z, alpha).clone()
    torch.manual_seed(77)
    draws = [[torch.empty(2, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2), device="cuda").normal_() for i in range(spec.num_layers)]
             for _ in range(3)]
    pipe._next_noise = None
    b1 = pipe.edit_device(z, alpha, noise=draws[0])
    b2 = pipe.edit_device(z, alpha, noise=draws[1])
    assert torch.equal(a1, b1) and torch.equal(a2, b2) and not torch.equal(a1, a2)
