"""CPU: identities the kernels rely on, latent/noise index maps, walks, z sampling."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from kernel_model import fused_forward_model
from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
from latent2im_b200.synthetic import synthetic_noise, synthetic_state_dict, synthetic_z
from oracle import GeneratorSpec, generator_forward_ref, mapping_ref, modulated_conv_ref
from oracle.walks import walk_linear_ref, walk_mlp_ref, walk_nonlinear_ref, z_sample_ref


def _setup(size=16, style_dim=32, n_mlp=2, batch=2, seed=0):
    spec = GeneratorSpec(size=size, style_dim=style_dim, n_mlp=n_mlp)
    shapes = {k: v.shape for k, v in Generator(size, style_dim, n_mlp).state_dict().items()}
    sd = {k: v.double() for k, v in synthetic_state_dict(shapes, seed).items()}
    lat = torch.randn(batch, spec.n_latent, style_dim, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    noise = synthetic_noise(spec.num_layers, batch)
    return spec, sd, lat, noise


def test_fused_algorithm_model_equals_oracle():
    spec, sd, lat, noise = _setup()
    ref, inter = generator_forward_ref(sd, lat, noise, spec, return_intermediates=True)
    img, acts, skips = fused_forward_model(sd, lat, noise, spec)
    assert (img - ref).abs().max() < 1e-11
    for name, a in acts.items():
        assert (a - inter[name]).abs().max() < 1e-11, name


def test_composite_upconv_model_equals_oracle():
    """conv_transpose2d(stride 2) + blur == per-phase 3x3 convs with the folded 6x6 kernel
    (pack_composite_weight_kernel), including the image borders."""
    spec, sd, lat, noise = _setup()
    ref, inter = generator_forward_ref(sd, lat, noise, spec, return_intermediates=True)
    img, acts, _ = fused_forward_model(sd, lat, noise, spec, composite=True)
    assert (img - ref).abs().max() < 1e-11
    for name, a in acts.items():
        assert (a - inter[name]).abs().max() < 1e-11, name


def test_rowfold_upconv_model_equals_oracle():
    """conv_transpose2d(stride 2) + blur == horizontally folded transposed-conv rows (pack_uprow_weight_kernel) followed
    by the vertical 4-tap FIR over rows u = -1 .. 2H+1 (conv_tc_uprow.cu), including the image borders."""
    spec, sd, lat, noise = _setup()
    ref, inter = generator_forward_ref(sd, lat, noise, spec, return_intermediates=True)
    img, acts, _ = fused_forward_model(sd, lat, noise, spec, composite="row")
    assert (img - ref).abs().max() < 1e-11
    for name, a in acts.items():
        assert (a - inter[name]).abs().max() < 1e-11, name


def test_modulation_is_input_scaling_identity():
    """weight modulation + grouped conv == scale input channels, shared conv, scale outputs by demod."""
    spec, sd, lat, _ = _setup()
    x = torch.randn(2, 512, 8, 8, dtype=torch.float64)
    ref = modulated_conv_ref(sd, "convs.1.conv", x, lat[:, 2])
    w = sd["convs.1.conv.weight"][0] / math.sqrt(512 * 9)
    s = lat[:, 2] @ (sd["convs.1.conv.modulation.weight"] / math.sqrt(32)).t() + sd["convs.1.conv.modulation.bias"]
    d = torch.rsqrt((s ** 2) @ (w ** 2).sum([2, 3]).t() + 1e-8)
    out = F.conv2d(x * s[:, :, None, None], w, padding=1) * d[:, :, None, None]
    assert (out - ref).abs().max() < 1e-11


def test_latent_index_map_is_respected():
    """Perturbing latent[:, i] must change the image iff layer i is consumed (all are) and the
    conv1 activation only for i == 0."""
    spec, sd, lat, noise = _setup()
    _, base = generator_forward_ref(sd, lat, noise, spec, return_intermediates=True)
    for i in range(spec.n_latent):
        lat2 = lat.clone()
        lat2[:, i] += 0.5
        _, inter = generator_forward_ref(sd, lat2, noise, spec, return_intermediates=True)
        changed_conv1 = (inter["conv1"] - base["conv1"]).abs().max() > 0
        assert bool(changed_conv1) == (i == 0)
        changed_rgb1 = (inter["to_rgb1"] - base["to_rgb1"]).abs().max() > 0
        assert bool(changed_rgb1) == (i in (0, 1))


def test_z_sampling_is_bit_exact_numpy_stream():
    a = z_sample_ref(5, seed=3)
    b = np.random.RandomState(3).randn(5, 512)
    assert a.dtype == np.float64 and np.array_equal(a, b) and np.array_equal(a, synthetic_z(5, 3))


def test_mapping_matches_manual():
    spec, sd, _, _ = _setup()
    z = torch.randn(3, 32, dtype=torch.float64)
    w = mapping_ref(sd, z, spec)
    x = z / torch.sqrt((z ** 2).mean(1, keepdim=True) + 1e-8)
    for i in (1, 2):
        x = x @ (sd[f"style.{i}.weight"] * 0.01 / math.sqrt(32)).t() + sd[f"style.{i}.bias"] * 0.01
        x = torch.where(x > 0, x, 0.2 * x) * math.sqrt(2)
    assert torch.allclose(w, x)


def test_walks():
    g = torch.Generator().manual_seed(0)
    ws = [torch.randn(3, 16, generator=g, dtype=torch.float64) for _ in range(4)]
    alpha = torch.randn(3, 2, generator=g, dtype=torch.float64)
    w = torch.randn(2, 4, 16, generator=g, dtype=torch.float64)
    out = walk_linear_ref(ws, alpha, w)
    for i in range(4):
        assert torch.allclose(out[i], ws[i] + alpha[:, :1] * w[0, i] + alpha[:, 1:] * w[1, i])
    out = walk_linear_ref(ws, alpha, w, layers=[1])
    assert torch.equal(out[0], ws[0]) and not torch.equal(out[1], ws[1])
    mlp = [(torch.randn(32, 16, generator=g, dtype=torch.float64), torch.randn(32, generator=g, dtype=torch.float64)),
           (torch.randn(16, 32, generator=g, dtype=torch.float64), torch.randn(16, generator=g, dtype=torch.float64))]
    out = walk_mlp_ref(ws, alpha, mlp)
    h = F.leaky_relu(ws[2] @ mlp[0][0].t() + mlp[0][1], 0.2) @ mlp[1][0].t() + mlp[1][1]
    assert torch.allclose(out[2], ws[2] + alpha[:, :1] * h)
    emb = (torch.randn(8, 10, generator=g, dtype=torch.float64), torch.randn(8, generator=g, dtype=torch.float64))
    mlp2 = [(torch.randn(32, 24, generator=g, dtype=torch.float64), torch.randn(32, generator=g, dtype=torch.float64)),
            (torch.randn(16, 32, generator=g, dtype=torch.float64), torch.randn(16, generator=g, dtype=torch.float64))]
    out = walk_nonlinear_ref(ws, alpha, emb, mlp2)
    assert torch.allclose((out[0] - ws[0]).norm(dim=1), torch.ones(3, dtype=torch.float64))
