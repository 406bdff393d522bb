"""CPU: pins the oracle against fixtures produced by the UNMODIFIED reference on a B200
(tests/golden/make_golden_ref_gpu.py -> tests/golden/ref_gpu_*.npz).  The reference ran in fp32
with TF32 disabled; the oracle runs in float64, so the gate is the fp32 round-off of the reference."""
import os

import numpy as np
import pytest
import torch

from latent2im_b200.synthetic import synthetic_noise, synthetic_state_dict, synthetic_z
from oracle import GeneratorSpec, generator_forward_ref, mapping_ref
from oracle.ops import fused_leaky_relu_bwd_ref, fused_leaky_relu_ref, upfirdn2d_ref
from oracle.walks import walk_linear_ref, walk_mlp_ref, walk_nonlinear_ref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not generated yet")
    return np.load(path)


def _t(a):
    return torch.from_numpy(np.asarray(a)).double()


def test_upfirdn2d_against_reference_op():
    z = _load("ref_gpu_ops.npz")
    n = len([k for k in z.files if k.startswith("upfirdn_") and k.endswith("_x")])
    assert n >= 6
    for i in range(n):
        up, down, p0, p1 = [int(v) for v in z[f"upfirdn_{i}_cfg"]]
        x = _t(z[f"upfirdn_{i}_x"]).requires_grad_(True)
        y = upfirdn2d_ref(x, _t(z[f"upfirdn_{i}_k"]), up=up, down=down, pad=(p0, p1))
        assert y.shape == z[f"upfirdn_{i}_y"].shape
        assert torch.allclose(y, _t(z[f"upfirdn_{i}_y"]), atol=2e-6, rtol=1e-5)
        (gx,) = torch.autograd.grad(y, x, _t(z[f"upfirdn_{i}_gy"]))
        assert torch.allclose(gx, _t(z[f"upfirdn_{i}_gx"]), atol=2e-6, rtol=1e-5)


def test_fused_leaky_relu_against_reference_op():
    z = _load("ref_gpu_ops.npz")
    for i in range(3):
        x, b = _t(z[f"flr_{i}_x"]), _t(z[f"flr_{i}_b"])
        y = fused_leaky_relu_ref(x, b)
        assert torch.allclose(y, _t(z[f"flr_{i}_y"]), atol=1e-6, rtol=1e-6)
        gx, gb = fused_leaky_relu_bwd_ref(_t(z[f"flr_{i}_gy"]), _t(z[f"flr_{i}_y"]))
        assert torch.allclose(gx, _t(z[f"flr_{i}_gx"]), atol=1e-6, rtol=1e-6)
        assert torch.allclose(gb, _t(z[f"flr_{i}_gb"]), atol=1e-4, rtol=1e-5)


@pytest.mark.parametrize("size", [16, 32, 64])
def test_generator_against_reference(size):
    from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
    z = _load("ref_gpu_generator.npz")
    size_, dim, n_mlp, batch = [int(v) for v in z[f"s{size}_cfg"]]
    spec = GeneratorSpec(size=size, style_dim=dim, n_mlp=n_mlp)
    shapes = {k: v.shape for k, v in Generator(size, dim, n_mlp).state_dict().items()}
    sd = {k: v.double() for k, v in synthetic_state_dict(shapes, seed=size).items()}
    # mapping network
    zz = torch.tensor(synthetic_z(batch, seed=size, dim_z=dim), dtype=torch.float32).double()
    w = mapping_ref(sd, zz, spec)
    assert torch.allclose(w, _t(z[f"s{size}_w"]), atol=5e-5, rtol=1e-4)
    # synthesis with explicit noise, then with the registered noise buffers
    lat = _t(z[f"s{size}_latent"]).requires_grad_(True)
    noise = synthetic_noise(spec.num_layers, batch, seed=2)
    img = generator_forward_ref(sd, lat, noise, spec)
    ref = _t(z[f"s{size}_image"])
    assert (img - ref).abs().max().item() <= 2e-4, (img - ref).abs().max().item()
    fixed = generator_forward_ref(sd, lat.detach(), [sd[f"noises.noise_{i}"] for i in range(spec.num_layers)], spec)
    assert (fixed - _t(z[f"s{size}_image_fixed_noise"])).abs().max().item() <= 2e-4
    # data gradient w.r.t. the W+ latent (the walk-training gradient path)
    (g,) = torch.autograd.grad((img * _t(z[f"s{size}_probe"])).sum(), lat)
    gref = _t(z[f"s{size}_grad_latent"])
    assert (g - gref).abs().max().item() <= 1e-3 * max(1.0, gref.abs().max().item())


def test_walks_against_reference_modules():
    z = _load("ref_gpu_walks.npz")
    ws = [_t(z["ws"])[:, i] for i in range(z["ws"].shape[1])]
    alpha = _t(z["alpha"])
    out = torch.stack(walk_linear_ref(ws, alpha, _t(z["linear_w"])), 1)
    assert torch.allclose(out, _t(z["linear_out"]), atol=1e-6)
    out = torch.stack(walk_linear_ref(ws, alpha, _t(z["linear_w"]), layers=[0, 3]), 1)
    assert torch.allclose(out, _t(z["linear_out_layers"]), atol=1e-6)
    mlp = [(_t(z[f"mlp_p{2 * i}"]), _t(z[f"mlp_p{2 * i + 1}"])) for i in range(3)]
    out = torch.stack(walk_mlp_ref(ws, alpha, mlp), 1)
    assert torch.allclose(out, _t(z["mlp_out"]), atol=1e-5)
    emb = (_t(z["nl_p0"]), _t(z["nl_p1"]))
    mlp2 = [(_t(z["nl_p2"]), _t(z["nl_p3"])), (_t(z["nl_p4"]), _t(z["nl_p5"]))]
    out = torch.stack(walk_nonlinear_ref(ws, alpha, emb, mlp2), 1)
    assert torch.allclose(out, _t(z["nl_out"]), atol=1e-5)
    out = torch.stack(walk_nonlinear_ref(ws, alpha, emb, mlp2, layers=[1, 2]), 1)
    assert torch.allclose(out, _t(z["nl_out_layers"]), atol=1e-5)


def test_oracle_matches_fullsize_reference_fixture_256():
    """cfg1 size: the float64 oracle against the image the unmodified reference produced on a B200 at 256 px (batch 2)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden_ref_gpu_fullsize import CASES, fullsize_inputs
    from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
    from latent2im_b200.synthetic import synthetic_state_dict
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_gpu_fullsize.npz"))
    tag, size, batch, seed = CASES[0]
    spec = GeneratorSpec(size=size)
    shapes = {k: v.shape for k, v in Generator(size, 512, 8).state_dict().items()}
    sd = {k: v.double() for k, v in synthetic_state_dict(shapes, seed).items()}
    w = torch.from_numpy(z[f"{tag}_w"])
    lat, noise, _ = fullsize_inputs(size, batch, seed, spec.n_latent, spec.num_layers, w)
    ref = generator_forward_ref(sd, lat.double(), [n.double() for n in noise], spec)
    err = (ref - torch.from_numpy(z[f"{tag}_image"]).double()).abs().max().item()
    assert err <= 1e-3, err
