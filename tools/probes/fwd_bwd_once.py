import sys, os
sys.path.insert(0, os.getcwd())
import torch
from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
from latent2im_b200.synthetic import load_synthetic, synthetic_z
B=16
gen = load_synthetic(Generator(1024, 512, 8), seed=0).cuda().eval()
gen.set_native(dtype=torch.bfloat16, max_batch=B)
z = torch.tensor(synthetic_z(B, 0), dtype=torch.float32).cuda()
w = gen.style(z)
lat = w[:, None, :].repeat(1, gen.n_latent, 1)
probe = torch.randn(B, 3, 1024, 1024, device="cuda")
for i in range(2):
    l = lat.clone().requires_grad_(True)
    img, _ = gen(l, input_is_latent=True)
    img.backward(probe)
torch.cuda.synchronize()
