"""Kernel timeline of the bench step (torch.profiler / CUPTI): per-kernel GPU time and the idle gaps between kernels over a few
back-to-back steps of the device-resident edit workload, without host synchronisation in between.
    python tools/trace_step.py [--steps 4] [--batch 32] [--size 1024] > gpurun_out/trace_step.txt"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("L2I_ALLOW_RANDOM_INIT", "1")

import torch
from torch.profiler import ProfilerActivity, profile


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=1024)
    args = ap.parse_args()
    from latent2im_b200 import _native as nt
    from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
    from latent2im_b200.graphs.stylegan_v2_real.transform_base import WalkLinearMultiW
    from latent2im_b200.pipeline import EditPipeline
    from latent2im_b200.synthetic import load_synthetic, synthetic_walk_w, synthetic_z
    dev = torch.device("cuda")
    nt.load()
    b = args.batch
    gen = load_synthetic(Generator(args.size, 512, 8), seed=0).to(dev).eval()
    gen.set_native(dtype=torch.bfloat16, max_batch=b)
    walk = WalkLinearMultiW(512, gen.log_size - 2, 1, ["Smiling"]).to(dev)
    with torch.no_grad():
        walk.w.copy_(synthetic_walk_w(1, gen.n_latent, 512, seed=0).to(dev))
    pipe = EditPipeline(gen, walk, b, n_attr=1, device=dev)
    z = torch.tensor(synthetic_z(b, seed=0), dtype=torch.float32, device=dev)
    alpha = torch.linspace(0, 1, b).reshape(b, 1).to(dev)
    for _ in range(5):
        pipe.edit_device(z, alpha)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(args.steps):
            pipe.edit_device(z, alpha)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    if not evs:
        print("no CUDA events recorded")
        return
    t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
    busy = sum(e.time_range.end - e.time_range.start for e in evs)
    # union of busy intervals (kernels of different streams overlap)
    cover, cur_s, cur_e = 0.0, None, None
    for e in evs:
        s, en = e.time_range.start, e.time_range.end
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                cover += cur_e - cur_s
            cur_s, cur_e = s, en
        else:
            cur_e = max(cur_e, en)
    cover += cur_e - cur_s
    span = t1 - t0
    print(f"{args.steps} steps: span {span / 1e3:.3f} ms = {span / 1e3 / args.steps:.3f} ms/step; GPU busy (union) {cover / 1e3:.3f} ms "
          f"({100 * cover / span:.1f} %), idle {100 * (1 - cover / span):.1f} %, sum of kernel times {busy / 1e3:.3f} ms, {len(evs)} device activities")
    by = collections.OrderedDict()
    for e in evs:
        k = e.name[:90]
        d = by.setdefault(k, [0, 0.0])
        d[0] += 1
        d[1] += e.time_range.end - e.time_range.start
    print(f"{'us/step':>10} {'n/step':>7}  kernel")
    for k, (n, t) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / args.steps:10.1f} {n / args.steps:7.1f}  {k}")
    # largest idle gaps
    gaps = []
    last_end = evs[0].time_range.end
    for a in evs[1:]:
        if a.time_range.start > last_end:
            gaps.append((a.time_range.start - last_end, a.name[:60]))
        last_end = max(last_end, a.time_range.end)
    gaps.sort(reverse=True)
    print("largest gaps (us, before kernel):", [(round(g, 1), n) for g, n in gaps[:12]])
    print(f"sum of gaps {sum(g for g, _ in gaps) / 1e3:.3f} ms over {args.steps} steps")


if __name__ == "__main__":
    main()
