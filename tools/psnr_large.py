"""PSNR of the bf16 tcgen05 path vs the fp32 CUDA-core path (itself gated against the oracle) at full size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
from latent2im_b200.synthetic import load_synthetic, synthetic_noise, synthetic_z

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 2
gen = load_synthetic(Generator(size, 512, 8), seed=0).cuda()
z = torch.tensor(synthetic_z(batch, 0), dtype=torch.float32).cuda()
w = gen.style(z)
lat = w[:, None, :].repeat(1, gen.n_latent, 1)
noise = [n.cuda() for n in synthetic_noise(gen.num_layers, batch)]
out = {}
for dt in (torch.float32, torch.bfloat16):
    gen.set_native(dtype=dt)
    out[dt], _ = gen(lat, input_is_latent=True, noise=noise)
torch.cuda.synchronize()
a, b = out[torch.float32].double(), out[torch.bfloat16].double()
mse = ((a - b) ** 2).mean().item()
print(f"size {size} batch {batch}: fp32 image std {a.std().item():.3f} range [{a.min().item():.2f},{a.max().item():.2f}] "
      f"bf16-vs-fp32 PSNR(peak-to-peak 2) {10 * torch.log10(torch.tensor(4.0 / mse)).item():.2f} dB  max-abs {(a - b).abs().max().item():.3e}")
