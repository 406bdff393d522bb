"""bf16-vs-fp32 PSNR at full size over several weight / latent seeds (accuracy margin of the split-bf16 threshold)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
from latent2im_b200.synthetic import load_synthetic, synthetic_noise, synthetic_z

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
gain = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
res, rel, rng = [], [], []
for seed in range(4):
    gen = load_synthetic(Generator(size, 512, 8), seed=seed, rgb_gain=gain).cuda()
    z = torch.tensor(synthetic_z(2, 10 + seed), dtype=torch.float32).cuda()
    lat = gen.style(z)[:, None, :].repeat(1, gen.n_latent, 1)
    noise = [n.cuda() for n in synthetic_noise(gen.num_layers, 2, seed=20 + seed)]
    out = {}
    for dt in (torch.float32, torch.bfloat16):
        gen.set_native(dtype=dt)
        out[dt], _ = gen(lat, input_is_latent=True, noise=noise)
    a, b = out[torch.float32].double(), out[torch.bfloat16].double()
    mse = ((a - b) ** 2).mean()
    res.append(10 * torch.log10(4.0 / mse).item())
    p2p = (a.max() - a.min()).item()
    rel.append(10 * torch.log10(p2p ** 2 / mse).item())
    rng.append(p2p)
    del gen
    torch.cuda.empty_cache()
print(f"size {size} rgb_gain {gain} L2I_SPLIT_RES={os.environ.get('L2I_SPLIT_RES', '64')}: PSNR(p2p 2) per seed " + " ".join(f"{r:.2f}" for r in res)
      + f"  min {min(res):.2f} | image p2p " + " ".join(f"{r:.1f}" for r in rng) + " | PSNR(own p2p) " + " ".join(f"{r:.2f}" for r in rel))
