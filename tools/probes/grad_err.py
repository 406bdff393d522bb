"""fp32 latent-gradient error vs the float64 oracle for the small configurations of tests/test_gpu_generator.py (debug probe)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch

from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
from latent2im_b200.synthetic import load_synthetic, synthetic_noise
from oracle import GeneratorSpec, generator_forward_ref

for size, dim, n_mlp, batch in [(8, 32, 1, 2), (16, 64, 2, 2), (32, 64, 1, 1), (16, 64, 2, 5)]:
    gen = load_synthetic(Generator(size, dim, n_mlp), seed=0)
    sd = {k: v.double() for k, v in gen.state_dict().items()}
    spec = GeneratorSpec(size=size, style_dim=dim, n_mlp=n_mlp)
    gen = gen.cuda()
    gen.set_native(dtype=torch.float32)
    lat = torch.randn(batch, spec.n_latent, dim, generator=torch.Generator().manual_seed(1))
    noise = synthetic_noise(spec.num_layers, batch)
    probe = torch.randn(batch, 3, size, size, generator=torch.Generator().manual_seed(9))
    lr = lat.double().requires_grad_(True)
    ref = generator_forward_ref(sd, lr, noise, spec)
    (gref,) = torch.autograd.grad((ref * probe.double()).sum(), lr)
    lc = lat.cuda().requires_grad_(True)
    img, _ = gen(lc, input_is_latent=True, noise=[n.cuda() for n in noise])
    (img * probe.cuda()).sum().backward()
    g = lc.grad.cpu().double()
    d = (g - gref).abs()
    i = d.flatten().argmax().item()
    print(f"BT={os.environ.get('L2I_LINEAR_BT', '1')} size {size} dim {dim} batch {batch}: image err {(img.detach().cpu().double() - ref.detach()).abs().max().item():.3e} "
          f"grad rel err {d.max().item() / gref.abs().max().item():.3e} at flat index {i} (layer {(i // dim) % spec.n_latent}), rel L2 {((g - gref).norm() / gref.norm()).item():.3e}")
