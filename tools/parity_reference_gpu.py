"""Full-size parity against the UNMODIFIED reference Generator running on the same B200 (its JIT-built ops + cuDNN,
TF32 off): fp32 kernels max-abs, bf16 kernels PSNR (peak-to-peak 2), same weights / latents / explicit noise.
    python oracle/stage_reference.py ; gpurun -- python tools/parity_reference_gpu.py
Measurement / test infrastructure only (baseline/_ref is git-ignored)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
os.environ.setdefault("TORCH_EXTENSIONS_DIR", os.path.join(REF, "_torch_ext"))
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

import torch

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def main():
    if not os.path.isdir(os.path.join(REF, "graphs")):
        print(json.dumps({"unavailable": "baseline/_ref not staged"}))
        return
    t0 = time.time()
    from graphs.stylegan_v2_real.networks import Generator as RefGenerator
    from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
    from latent2im_b200.synthetic import synthetic_noise, synthetic_state_dict, synthetic_z
    dev = torch.device("cuda")
    out = {"reference_ops_build_s": round(time.time() - t0, 1), "tf32": False, "cases": []}
    for size, batch, gain, seed in ((256, 4, 1.0, 0), (256, 4, 0.25, 1), (1024, 2, 0.25, 0), (1024, 2, 1.0, 2)):
        ref = RefGenerator(size, 512, 8).to(dev).eval()
        sd = synthetic_state_dict({k: v.shape for k, v in ref.state_dict().items()}, seed, rgb_gain=gain)
        ref.load_state_dict({k: v.to(dev) for k, v in sd.items()}, strict=False)
        gen = Generator(size, 512, 8)
        gen.load_state_dict(sd, strict=False)
        gen = gen.to(dev).eval()
        z = torch.tensor(synthetic_z(batch, 10 + seed), dtype=torch.float32, device=dev)
        noise = [n.to(dev) for n in synthetic_noise(ref.num_layers, batch, seed=20 + seed)]
        with torch.no_grad():
            w_ref = ref.style(z)
            w = gen.style(z)
            lat = w_ref[:, None, :].repeat(1, ref.n_latent, 1)
            img_ref, _ = ref(lat, input_is_latent=True, noise=noise)
            res = {"size": size, "batch": batch, "rgb_gain": gain, "seed": seed,
                   "mapping_max_abs": (w - w_ref).abs().max().item(),
                   "image_peak_to_peak": (img_ref.max() - img_ref.min()).item()}
            for name, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
                gen.set_native(dtype=dt, max_batch=batch)
                img, _ = gen(lat, input_is_latent=True, noise=noise)
                d = (img.double() - img_ref.double())
                res[name + "_max_abs"] = d.abs().max().item()
                res[name + "_psnr_p2p2_db"] = (10 * torch.log10(4.0 / (d ** 2).mean())).item()
        out["cases"].append(res)
        del ref, gen
        torch.cuda.empty_cache()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
