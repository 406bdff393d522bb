"""Full-size golden fixtures (BASELINE.json cfg1 = 256 px, cfg2 = 1024 px) from the UNMODIFIED reference on a CUDA device.

    python oracle/stage_reference.py                               # build container, once
    gpurun -- python tests/golden/make_golden_ref_gpu_fullsize.py
    cp gpurun_out/golden/ref_gpu_fullsize.npz tests/golden/        # commit

Recipe = SURVEY 8d, UNSCALED (rgb_gain 1.0): weights ``synthetic_state_dict(seed)``, z ``RandomState``, explicit noise.
Inputs are rebuilt from seeds by the tests (``fullsize_inputs`` below is imported by them), only outputs are stored:
  * 256 px, batch 2: the whole fp32 image + the latent gradient of ``sum(image * probe)``;
  * 1024 px, batch 1: six 128 x 128 fp32 crops, the 4 x 4 average-pooled image (pins every pixel in aggregate),
    per-channel moments, and the latent gradient.
TF32 is disabled: the fixtures are the reference's fp32 arithmetic (its JIT-built upfirdn2d / fused_bias_act + cuDNN).
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "baseline", "_ref")

CASES = (("s256", 256, 2, 0), ("s1024", 1024, 1, 2))       # tag, size, batch, seed
CROPS_1024 = ((0, 0), (0, 896), (896, 0), (896, 896), (448, 448), (301, 611))   # (y, x) of the 128 x 128 crops
CROP = 128


def fullsize_inputs(size, batch, seed, n_latent, num_layers, w):
    """Latent / noise / probe of one case, rebuilt from seeds (used by the generator script AND by the tests).
    ``w``: [batch, 512] mapping output on any device; returns tensors on ``w.device``."""
    import torch
    from latent2im_b200.synthetic import synthetic_noise
    g = torch.Generator().manual_seed(1000 + size + seed)
    lat = w[:, None, :] + 0.3 * torch.randn(batch, n_latent, w.shape[1], generator=g).to(w.device)
    noise = [n.to(w.device) for n in synthetic_noise(num_layers, batch, seed=20 + seed)]
    probe = torch.randn(batch, 3, size, size, generator=g).to(w.device)
    return lat.detach(), noise, probe


def main(out_dir):
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("TORCH_EXTENSIONS_DIR", os.path.join(REF, "_torch_ext"))
    sys.path.insert(0, REF)
    sys.path.insert(1, ROOT)
    import numpy as np
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    os.makedirs(out_dir, exist_ok=True)
    dev = torch.device("cuda")
    t0 = time.time()
    from graphs.stylegan_v2_real.networks import Generator as RefGenerator
    print(f"reference imported (ops built) in {time.time() - t0:.1f}s", flush=True)
    from latent2im_b200.synthetic import synthetic_state_dict, synthetic_z

    out = {}
    for tag, size, batch, seed in CASES:
        ref = RefGenerator(size, 512, 8)
        syn = synthetic_state_dict({k: v.shape for k, v in ref.state_dict().items()}, seed=seed)
        missing = ref.load_state_dict(syn, strict=False)
        assert not missing.unexpected_keys, missing
        ref = ref.to(dev).eval()
        z = torch.tensor(synthetic_z(batch, 10 + seed), dtype=torch.float32, device=dev)
        with torch.no_grad():
            w = ref.style(z)
        lat, noise, probe = fullsize_inputs(size, batch, seed, ref.n_latent, ref.num_layers, w)
        lat.requires_grad_(True)
        img, _ = ref(lat, input_is_latent=True, noise=noise)
        (glat,) = torch.autograd.grad((img * probe).sum(), lat)
        img = img.detach()
        out[f"{tag}_cfg"] = np.array([size, batch, seed])
        out[f"{tag}_w"] = w.cpu().numpy()
        out[f"{tag}_grad_latent"] = glat.cpu().numpy()
        out[f"{tag}_moments"] = torch.stack([img.mean((0, 2, 3)), img.std((0, 2, 3)), img.amin((0, 2, 3)), img.amax((0, 2, 3))]).cpu().numpy()
        if size <= 256:
            out[f"{tag}_image"] = img.cpu().numpy()
        else:
            out[f"{tag}_pooled4"] = torch.nn.functional.avg_pool2d(img.double(), 4).float().cpu().numpy()
            out[f"{tag}_crops"] = np.stack([img[:, :, y:y + CROP, x:x + CROP].cpu().numpy() for (y, x) in CROPS_1024])
        print(f"{tag}: image range [{img.min().item():.3f}, {img.max().item():.3f}] std {img.std().item():.3f} "
              f"|grad| max {glat.abs().max().item():.3e}", flush=True)
        del ref, img, glat
        torch.cuda.empty_cache()
    np.savez_compressed(os.path.join(out_dir, "ref_gpu_fullsize.npz"), **out)
    print("done", flush=True)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
