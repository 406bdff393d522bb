import os, sys, time, cProfile, pstats
sys.path.insert(0, os.getcwd())
os.environ.setdefault("L2I_ALLOW_RANDOM_INIT", "1")
import torch
from latent2im_b200 import _native as nt
from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
from latent2im_b200.graphs.stylegan_v2_real.transform_base import WalkLinearMultiW
from latent2im_b200.pipeline import EditPipeline
from latent2im_b200.synthetic import load_synthetic, synthetic_walk_w, synthetic_z
dev = torch.device("cuda"); nt.load()
b, size = 4, 256
gen = load_synthetic(Generator(size, 512, 8), seed=0).to(dev).eval()
gen.set_native(dtype=torch.bfloat16, max_batch=b)
walk = WalkLinearMultiW(512, gen.log_size - 2, 1, ["Smiling"]).to(dev)
pipe = EditPipeline(gen, walk, b, n_attr=1, device=dev)
z = torch.tensor(synthetic_z(b, 0), dtype=torch.float32, device=dev); alpha = torch.linspace(0, 1, b).reshape(b, 1).to(dev)
for _ in range(20): pipe.edit_device(z, alpha, want_uint8=True)
torch.cuda.synchronize()
n = 200
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(n): pipe.edit_device(z, alpha, want_uint8=True)
e1.record(); t_host = time.perf_counter() - t0
torch.cuda.synchronize(); t_all = time.perf_counter() - t0
print(f"256px B=4: host enqueue {1e3*t_host/n:.3f} ms/call, wall {1e3*t_all/n:.3f} ms/call, GPU span {e0.elapsed_time(e1)/n:.3f} ms/call")
pr = cProfile.Profile(); pr.enable()
for _ in range(n): pipe.edit_device(z, alpha, want_uint8=True)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
