"""Times the UNMODIFIED reference Generator (its JIT-built upfirdn2d / fused ops + cuDNN grouped convs) on the
same B200, same workload as bench.py's device-resident arm, next to this repository's kernels.

    python oracle/stage_reference.py                          # build container: stage sources into baseline/_ref
    gpurun -- python tools/bench_reference_gpu.py [--size 1024] [--batch 32] [--steps 5]

The reference has no bf16 path (AT_DISPATCH_FLOATING_TYPES_AND_HALF) - it runs fp32, timed with TF32 both on
(the PyTorch default for cuDNN convs) and off, and fp16 where its ops allow.  Test/measurement infrastructure only.
"""
import argparse
import json
import os
os.environ.setdefault("L2I_ALLOW_RANDOM_INIT", "1")   # synthetic-weight benchmark
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
os.environ.setdefault("TORCH_EXTENSIONS_DIR", os.path.join(REF, "_torch_ext"))
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

import torch


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--ref-only", action="store_true", help="time only the reference (bench.py's ref_gpu_baseline key)")
    a = ap.parse_args()
    if not os.path.isdir(os.path.join(REF, "graphs")):
        print(json.dumps({"impl": "reference_gpu", "unavailable": "baseline/_ref not staged (run oracle/stage_reference.py)"}))
        return
    t0 = time.time()
    from graphs.stylegan_v2_real.networks import Generator as RefGenerator
    build_s = time.time() - t0
    from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
    from latent2im_b200.synthetic import synthetic_state_dict, synthetic_z

    dev = torch.device("cuda")
    ref = RefGenerator(a.size, 512, 8).to(dev).eval()
    sd = synthetic_state_dict({k: v.shape for k, v in ref.state_dict().items()}, 0)
    ref.load_state_dict({k: v.to(dev) for k, v in sd.items()}, strict=False)
    z = torch.tensor(synthetic_z(a.batch, 0), dtype=torch.float32, device=dev)
    out = {"impl": "reference_gpu", "size": a.size, "batch": a.batch, "ops_build_s": round(build_s, 1)}
    with torch.no_grad():
        w = ref.style(z)
        lat = w[:, None, :].repeat(1, ref.n_latent, 1)

        def step():
            ref(lat, input_is_latent=True)

        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            try:
                ms = timed(step, a.steps, a.warmup)
                out["fp32_tf32_on" if tf32 else "fp32_tf32_off"] = {"ms_per_step": ms, "images_per_s": a.batch / ms * 1e3}
            except RuntimeError as e:  # e.g. out of memory at this batch
                out["fp32_tf32_on" if tf32 else "fp32_tf32_off"] = {"error": str(e)[:200]}
                torch.cuda.empty_cache()
        try:
            if a.ref_only:
                raise RuntimeError("skipped (--ref-only)")
            ref16, lat16 = ref.half(), lat.half()
            ms = timed(lambda: ref16(lat16, input_is_latent=True), a.steps, a.warmup)
            out["fp16"] = {"ms_per_step": ms, "images_per_s": a.batch / ms * 1e3}
        except RuntimeError as e:
            out["fp16"] = {"error": str(e)[:200]}
        del ref
        torch.cuda.empty_cache()

        if a.ref_only:
            print(json.dumps(out), flush=True)
            return
        # this repository, same weights / latents / fresh noise, device-resident
        gen = Generator(a.size, 512, 8)
        gen.load_state_dict(sd, strict=False)
        gen = gen.to(dev).eval()
        for name, dt in (("ours_bf16", torch.bfloat16), ("ours_fp32", torch.float32)):
            gen.set_native(dtype=dt, max_batch=a.batch)
            ms = timed(lambda: gen(lat.float(), input_is_latent=True), a.steps if dt == torch.float32 else 10 * a.steps, a.warmup)
            out[name] = {"ms_per_step": ms, "images_per_s": a.batch / ms * 1e3}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
