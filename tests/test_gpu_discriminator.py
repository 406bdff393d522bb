"""Discriminator on the sm_100a ops (blur = l2i_upfirdn2d, bias + leaky-relu = l2i_fused_bias_act; convs on cuDNN, TF32 off)
against the fixture of the unmodified reference module (tests/golden/ref_gpu_discriminator.npz): logits and the
gradient with respect to the image, i.e. forward and backward of both native ops inside the real network."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _image(batch, size, seed):
    g = torch.Generator().manual_seed(seed)
    return 0.5 * torch.randn(batch, 3, size, size, generator=g, dtype=torch.float32)


@pytest.fixture
def no_tf32():
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_discriminator_matches_reference_fixture(no_tf32):
    from latent2im_b200.graphs.stylegan_v2_real.networks import Discriminator
    from latent2im_b200.synthetic import synthetic_discriminator_state_dict
    fx = np.load(os.path.join(GOLD, "ref_gpu_discriminator.npz"))
    for case in json.loads(str(fx["cases"])):
        size, cm, batch, seed = case["size"], case["cm"], case["batch"], case["seed"]
        d = Discriminator(size, channel_multiplier=cm)
        sd = synthetic_discriminator_state_dict({k: v.shape for k, v in d.state_dict().items()}, seed)
        d.load_state_dict(sd, strict=False)
        d = d.cuda().eval()
        x = _image(batch, size, 100 + seed).cuda().requires_grad_(True)
        out = d(x)
        ref = torch.tensor(fx[case["name"] + "_logits"]).cuda()
        assert out.shape == ref.shape
        assert float((out.detach() - ref).abs().max()) <= 2e-4 * max(float(ref.abs().max()), 1.0), case
        if case["grad"]:
            coef = torch.tensor(fx[case["name"] + "_coef"]).cuda()
            (out * coef).sum().backward()
            ref_g = torch.tensor(fx[case["name"] + "_grad"]).cuda()
            assert float((x.grad - ref_g).abs().max()) <= 2e-4 * float(ref_g.abs().max()), case


def test_gan_term_flows_to_the_walk(no_tf32):
    """Reference transform_base.py:455-462: BCE-with-logits of D(G(w')) against ones; its gradient must reach W+."""
    from latent2im_b200.graphs.stylegan_v2_real.networks import Discriminator, Generator
    from latent2im_b200.synthetic import load_synthetic, synthetic_discriminator_state_dict, synthetic_noise, synthetic_z
    size = 32
    gen = load_synthetic(Generator(size, 512, 2), seed=0, rgb_gain=0.25).cuda().eval()
    gen.set_native(dtype=torch.float32, max_batch=4)
    d = Discriminator(size)
    d.load_state_dict(synthetic_discriminator_state_dict({k: v.shape for k, v in d.state_dict().items()}, 3), strict=False)
    d = d.cuda().eval()
    z = torch.tensor(synthetic_z(4, 0), dtype=torch.float32).cuda()
    lat = gen.style(z)[:, None, :].repeat(1, gen.n_latent, 1).clone().requires_grad_(True)
    img, _ = gen(lat, input_is_latent=True, noise=[n.cuda() for n in synthetic_noise(gen.num_layers, 4)])
    logit = d(img)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logit, torch.ones_like(logit))
    loss.backward()
    assert torch.isfinite(loss) and lat.grad is not None and torch.isfinite(lat.grad).all() and float(lat.grad.abs().max()) > 0


def test_train_with_gan_and_content_terms(tmp_path, monkeypatch):
    """train.py without --no_gan_loss / --no_content_loss: loss = 10 reg + 0.05 content + 0.05 gan (transform_base.py:470-483)."""
    import math
    import re

    import train
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    attr = os.path.join(root, "latent2im_b200", "dataset", "attributes_celeba.txt")
    monkeypatch.setenv("L2I_G_PATH", "/nonexistent")
    monkeypatch.setenv("L2I_REG_PATH", "/nonexistent")
    out = train.main(["--model", "stylegan_v2_real", "--transform", "face", "--num_samples", "8", "--learning_rate", "1e-3",
                      "--latent", "w", "--walk_type", "linear", "--loss", "l2", "--attrList", "Smiling", "--attrPath", attr,
                      "--models_dir", str(tmp_path), "--overwrite_config", "--size", "32", "--batch_size", "4", "--dtype", "fp32",
                      "--epochs", "1", "--max_iters", "2", "--log_every", "1"])
    assert os.path.exists(os.path.join(out, "model_w_1_final_walk_module.ckpt"))
    losses = [float(m) for m in re.findall(r"lss, alpha: [^,]+, \d+, \d+, ([^,]+),", open(os.path.join(out, "log.txt")).read())]
    assert len(losses) == 2 and all(math.isfinite(v) and v > 0 for v in losses)
