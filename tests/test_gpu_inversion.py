"""GPU: W+ inversion (reference BP.py) through the native forward + data-gradient kernels recovers an image the
generator itself produced, and the first step's gradient equals autograd through the float64 oracle."""
import pytest
import torch

from latent2im_b200.synthetic import load_synthetic, synthetic_noise
from oracle import GeneratorSpec, generator_forward_ref

pytestmark = pytest.mark.gpu


def test_inversion_reduces_reconstruction_error():
    from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
    from latent2im_b200.inversion import invert, reconstruction_loss
    size, dim, batch = 32, 64, 2
    spec = GeneratorSpec(size=size, style_dim=dim, n_mlp=2)
    gen = load_synthetic(Generator(size, dim, 2), seed=0, rgb_gain=0.25)
    sd = {k: v.double() for k, v in gen.state_dict().items()}
    gen = gen.cuda().eval()
    gen.set_native(dtype=torch.float32, max_batch=batch)
    noise = [n.cuda() for n in synthetic_noise(spec.num_layers, batch)]
    w_true = 0.5 * torch.randn(batch, spec.n_latent, dim, generator=torch.Generator().manual_seed(4)).cuda()
    with torch.no_grad():
        target, _ = gen(w_true, input_is_latent=True, noise=noise)
    mean = torch.zeros(1, dim, device="cuda")
    w, losses = invert(gen, target, n_loops=60, lr=5e-2, noise=noise, mean_latent=mean)
    assert w.shape == (batch, spec.n_latent, dim)
    assert losses[-1].item() < 0.7 * losses[0].item()

    # gradient of the first step vs the oracle
    w0 = mean.reshape(1, 1, -1).repeat(batch, spec.n_latent, 1).clone().requires_grad_(True)
    out, _ = gen(w0, input_is_latent=True, noise=noise)
    reconstruction_loss(out, target).sum().backward()
    wr = w0.detach().cpu().double().requires_grad_(True)
    ref = generator_forward_ref(sd, wr, [n.cpu() for n in noise], spec)
    reconstruction_loss(ref, target.cpu().double()).sum().backward()
    scale = wr.grad.abs().max().item()
    assert (w0.grad.cpu().double() - wr.grad).abs().max().item() <= 2e-3 * scale
