"""One bf16 forward + backward at 512 px (channel multiplier 1: every special tcgen05 kernel is on the path) and one bf16 forward at
256 px (channel multiplier 2: the 256 -> 128 row-marching up-conv and the A-resident 128 -> 128 layer) for compute-sanitizer
(--tool memcheck | racecheck | initcheck | synccheck)."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
from latent2im_b200.synthetic import load_synthetic, synthetic_noise, synthetic_z
gen = load_synthetic(Generator(512, 512, 2, channel_multiplier=1), seed=0, rgb_gain=0.25).cuda().eval()
gen.set_native(dtype=torch.bfloat16, max_batch=2)
z = torch.tensor(synthetic_z(2, 0), dtype=torch.float32).cuda()
lat = gen.style(z)[:, None, :].repeat(1, gen.n_latent, 1)
noise = [n.cuda() for n in synthetic_noise(gen.num_layers, 2)]
with torch.no_grad():
    img, u8 = gen.synthesize(lat, noise=noise, want_uint8=True, want_float=True)
    u8b = gen.synthesize(lat, noise=noise, want_uint8=True, want_float=False)
l = lat.clone().requires_grad_(True)
out, _ = gen(l, input_is_latent=True, noise=noise)
out.sum().backward()
torch.cuda.synchronize()
print("forward/backward done", float(img.abs().mean()), float(l.grad.abs().mean()), bool(torch.equal(u8, u8b)))
gen2 = load_synthetic(Generator(256, 512, 2), seed=1, rgb_gain=0.25).cuda().eval()
gen2.set_native(dtype=torch.bfloat16, max_batch=1)
z2 = torch.tensor(synthetic_z(1, 1), dtype=torch.float32).cuda()
with torch.no_grad():
    lat2 = gen2.style(z2)[:, None, :].repeat(1, gen2.n_latent, 1)
    img2, _ = gen2(lat2, input_is_latent=True, noise=[n.cuda() for n in synthetic_noise(gen2.num_layers, 1)])
torch.cuda.synchronize()
print("256 px forward done", float(img2.abs().mean()))
