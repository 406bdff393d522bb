"""Bring-up aid for the halo-resident tcgen05 kernel: saves / compares the 1024px image for one env setting."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
from latent2im_b200.synthetic import load_synthetic, synthetic_noise, synthetic_z

size, batch = int(sys.argv[1]), int(sys.argv[2])
tag = sys.argv[3]
gen = load_synthetic(Generator(size, 512, 8), seed=0).cuda()
gen.set_native(dtype=torch.bfloat16)
z = torch.tensor(synthetic_z(batch, 0), dtype=torch.float32).cuda()
w = gen.style(z)
lat = w[:, None, :].repeat(1, gen.n_latent, 1)
noise = [n.cuda() for n in synthetic_noise(gen.num_layers, batch)]
img, _ = gen(lat, input_is_latent=True, noise=noise)
torch.cuda.synchronize()
path = os.path.join(ROOT, "gpurun_out", f"halo_{tag}.pt")
ref_path = os.path.join(ROOT, "gpurun_out", "halo_ref.pt")
if tag == "ref":
    torch.save(img.cpu(), ref_path)
    print("saved reference image (halo kernel off)", img.std().item())
else:
    ref = torch.load(ref_path).double()
    a = img.cpu().double()
    mse = ((a - ref) ** 2).mean().item()
    import math
    print(f"{tag}: vs halo-off image: PSNR {10 * math.log10(4.0 / max(mse, 1e-30)):.2f} dB  max-abs {(a - ref).abs().max().item():.3e}  nan={torch.isnan(a).any().item()}")
