import os, sys
sys.path.insert(0, os.getcwd())
import torch
from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
from latent2im_b200.synthetic import load_synthetic, synthetic_z
gen = load_synthetic(Generator(64, 512, 8), seed=0).cuda().eval()
gen.set_native(dtype=torch.bfloat16, max_batch=32)
z = torch.tensor(synthetic_z(32, 0), dtype=torch.float32).cuda()
for _ in range(5): gen.style(z)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200): gen.style(z)
e1.record(); torch.cuda.synchronize()
print("mapping B=32: %.1f us per call" % (e0.elapsed_time(e1) / 200 * 1e3))
