"""Discriminator: checkpoint layout, module wiring and the oracle's pin - all on CPU."""
import json
import os

import numpy as np
import pytest
import torch

from latent2im_b200.graphs.stylegan_v2_real.networks import Discriminator
from latent2im_b200.synthetic import synthetic_discriminator_state_dict
from oracle.discriminator import discriminator_forward_ref
from oracle.ops import fused_leaky_relu_ref, upfirdn2d_ref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _image(batch, size, seed):
    g = torch.Generator().manual_seed(seed)
    return 0.5 * torch.randn(batch, 3, size, size, generator=g, dtype=torch.float32)


@pytest.mark.parametrize("tag", ["32x2", "64x1", "256x2", "1024x2"])
def test_state_dict_layout_matches_reference(tag):
    """Key names, order and shapes equal the reference Discriminator's (tests/golden/make_ref_keys.py)."""
    ref = json.load(open(os.path.join(GOLD, "ref_discriminator_keys.json")))[tag]
    size, cm = (int(v) for v in tag.split("x"))
    with torch.device("meta"):
        d = Discriminator(size, channel_multiplier=cm)
    assert [(k, list(v.shape)) for k, v in d.state_dict().items()] == [(k, list(s)) for k, s in ref]


def test_module_wiring_equals_oracle_with_oracle_ops():
    """The product module run with the oracle's restatements of its two native ops is the oracle network
    (float64; the native ops themselves are covered by tests/test_gpu_ops.py)."""
    size, batch = 16, 8
    d = Discriminator(size, channel_multiplier=1).double()
    sd = synthetic_discriminator_state_dict({k: v.shape for k, v in d.state_dict().items()}, seed=4)
    d.load_state_dict(sd, strict=False)
    x = _image(batch, size, 5).double()
    got = d.run(x, upfirdn2d_ref, fused_leaky_relu_ref)
    want = discriminator_forward_ref({k: v.double() for k, v in d.state_dict().items()}, x, size)
    assert got.shape == (batch, 1)
    assert torch.allclose(got, want, rtol=1e-10, atol=1e-10)


def test_product_forward_needs_cuda():
    d = Discriminator(16, channel_multiplier=1)
    with pytest.raises(RuntimeError):
        d(_image(2, 16, 0))


def test_oracle_matches_reference_gpu_fixture():
    """Logits and input gradient of the UNMODIFIED reference module on a B200 (fp32, TF32 off) vs the float64 oracle."""
    path = os.path.join(GOLD, "ref_gpu_discriminator.npz")
    fx = np.load(path)
    for case in json.loads(str(fx["cases"])):
        size, cm, batch, seed = case["size"], case["cm"], case["batch"], case["seed"]
        shapes = {k: tuple(s) for k, s in json.load(open(os.path.join(GOLD, "ref_discriminator_keys.json")))[f"{size}x{cm}"]}
        sd = {k: v.double() for k, v in synthetic_discriminator_state_dict(shapes, seed).items()}
        x = _image(batch, size, 100 + seed).double().requires_grad_(True)
        out = discriminator_forward_ref(sd, x, size)
        ref_out = torch.tensor(fx[case["name"] + "_logits"]).double()
        scale = float(ref_out.abs().max())
        assert float((out.detach() - ref_out).abs().max()) <= 2e-4 * max(scale, 1.0), case
        if case["grad"]:
            coef = torch.tensor(fx[case["name"] + "_coef"]).double()
            (out * coef).sum().backward()
            ref_g = torch.tensor(fx[case["name"] + "_grad"]).double()
            # the fixture's backward ran in fp32 through eleven 512-channel convs: measured 1.3e-3 of the peak (rms 2e-4)
            err = (x.grad - ref_g).abs()
            assert float(err.max()) <= 5e-3 * float(ref_g.abs().max()), case
            assert float(err.pow(2).mean().sqrt()) <= 1e-3 * float(ref_g.pow(2).mean().sqrt()), case
