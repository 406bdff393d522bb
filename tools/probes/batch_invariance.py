"""Diagnostic: image i of a large batch vs the same latent run alone; repeated-run determinism (per kernel switch)."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
from latent2im_b200.synthetic import load_synthetic, synthetic_noise, synthetic_z
size, batch = 1024, int(os.environ.get("PB", "40"))
gen = load_synthetic(Generator(size, 512, 2), seed=1, rgb_gain=0.25).cuda()
gen.set_native(dtype=torch.bfloat16, max_batch=batch)
z = torch.tensor(synthetic_z(batch, 3), dtype=torch.float32).cuda()
lat = gen.style(z)[:, None, :].repeat(1, gen.n_latent, 1)
noise = [n.cuda() for n in synthetic_noise(gen.num_layers, batch)]
with torch.no_grad():
    big, _ = gen(lat, input_is_latent=True, noise=noise)
    big = big.clone()
    bad = 0
    for rep in range(int(os.environ.get("PREP", "6"))):
        big2, _ = gen(lat, input_is_latent=True, noise=noise)
        d = (big - big2).abs()
        if float(d.max()) != 0.0:
            bad += 1
            print("  repeat", rep, "differs: n =", int((d > 0).sum()), "max", float(d.max()), torch.nonzero(d > 0)[0].tolist())
    print("repeat runs differing:", bad)
    for i in (0, 1, 17, batch - 1):
        one, _ = gen(lat[i:i + 1], input_is_latent=True, noise=[n[i:i + 1] for n in noise])
        d = (one[0] - big[i]).abs()
        nz = torch.nonzero(d > 0)
        msg = ""
        if nz.numel():
            ys, xs = nz[:, 1], nz[:, 2]
            msg = f"  n={nz.shape[0]} y[{int(ys.min())},{int(ys.max())}] x[{int(xs.min())},{int(xs.max())}] first={nz[0].tolist()}"
        print(i, "max diff", float(d.max()), msg)
