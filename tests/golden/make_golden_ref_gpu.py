"""Generates the golden fixtures by running the UNMODIFIED reference on a CUDA device.

Run on the B200 box (the reference's path is CUDA-only):

    python oracle/stage_reference.py                   # in the build container, once
    gpurun -- python tests/golden/make_golden_ref_gpu.py
    cp gpurun_out/golden/ref_gpu_*.npz tests/golden/    # commit

Inputs are rebuilt from seeds by ``latent2im_b200.synthetic`` so only outputs (a few hundred KB)
are stored.  TF32 is disabled: the fixtures are the reference's fp32 arithmetic.
"""
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "baseline", "_ref")
os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
os.environ.setdefault("TORCH_EXTENSIONS_DIR", os.path.join(REF, "_torch_ext"))
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

import numpy as np
import torch

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _stub_missing_modules():
    """transform_base.py imports helper packages that are not part of the hot path and not
    installed here (easydict, the repo's utils.image -> cv2).  Empty stand-ins let the file import;
    none of them is touched by the walk modules."""
    for name in ("easydict", "utils", "utils.image"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            sys.modules[name] = m
    sys.modules["easydict"].EasyDict = dict
    sys.modules["utils"].image = sys.modules["utils.image"]


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    dev = torch.device("cuda")
    t0 = time.time()
    from graphs.stylegan_v2_real.networks import Generator as RefGenerator
    from graphs.stylegan_v2_real.op import fused_leaky_relu as ref_flr, upfirdn2d as ref_upfirdn2d
    print(f"reference imported (ops built) in {time.time() - t0:.1f}s", flush=True)
    from latent2im_b200.synthetic import synthetic_noise, synthetic_state_dict, synthetic_walk_w, synthetic_z

    # ---- native ops ---------------------------------------------------------------------------
    ops = {}
    g = torch.Generator().manual_seed(11)
    cases = [  # (shape, kernel taps, up, down, pad)  - the modes the generator uses + odd sizes
        ((2, 3, 9, 9), [1, 3, 3, 1], 1, 1, (1, 1)),      # Blur after up-conv (mode 1)
        ((2, 3, 8, 8), [1, 3, 3, 1], 2, 1, (2, 1)),      # Upsample (mode 3)
        ((2, 3, 16, 16), [1, 3, 3, 1], 1, 2, (1, 1)),    # Upsample backward / Downsample (mode 5)
        ((1, 2, 7, 5), [1, 3, 3, 1], 1, 1, (2, 2)),      # ragged
        ((1, 2, 33, 70), [1, 3, 3, 1], 2, 1, (2, 1)),    # crosses the reference's 16x64 tiles
        ((1, 1, 12, 12), [1, 2, 1], 1, 1, (1, 1)),       # 3-tap (mode 2)
    ]
    for i, (shape, taps, up, down, pad) in enumerate(cases):
        x = torch.randn(shape, generator=g)
        k = torch.tensor(taps, dtype=torch.float32)
        k = k[None, :] * k[:, None]
        k = k / k.sum() * (up ** 2)
        xg = x.to(dev).requires_grad_(True)
        y = ref_upfirdn2d(xg, k.to(dev), up=up, down=down, pad=pad)
        gy = torch.randn(y.shape, generator=g)
        (gx,) = torch.autograd.grad(y, xg, gy.to(dev))
        ops[f"upfirdn_{i}_x"], ops[f"upfirdn_{i}_k"] = x.numpy(), k.numpy()
        ops[f"upfirdn_{i}_cfg"] = np.array([up, down, pad[0], pad[1]])
        ops[f"upfirdn_{i}_y"], ops[f"upfirdn_{i}_gy"], ops[f"upfirdn_{i}_gx"] = y.detach().cpu().numpy(), gy.numpy(), gx.cpu().numpy()
    for i, shape in enumerate([(2, 8, 5, 5), (3, 16), (1, 4, 33, 17)]):
        x = torch.randn(shape, generator=g)
        b = torch.randn(shape[1], generator=g)
        xg, bg = x.to(dev).requires_grad_(True), b.to(dev).requires_grad_(True)
        y = ref_flr(xg, bg)
        gy = torch.randn(shape, generator=g)
        gx, gb = torch.autograd.grad(y, (xg, bg), gy.to(dev))
        ops[f"flr_{i}_x"], ops[f"flr_{i}_b"], ops[f"flr_{i}_y"] = x.numpy(), b.numpy(), y.detach().cpu().numpy()
        ops[f"flr_{i}_gy"], ops[f"flr_{i}_gx"], ops[f"flr_{i}_gb"] = gy.numpy(), gx.cpu().numpy(), gb.cpu().numpy()
    np.savez_compressed(os.path.join(out_dir, "ref_gpu_ops.npz"), **ops)
    print("ops fixtures written", flush=True)

    # ---- generator ----------------------------------------------------------------------------
    gen_out = {}
    for size, style_dim, n_mlp, batch in [(16, 512, 8, 2), (32, 64, 2, 3), (64, 512, 8, 1)]:
        ref = RefGenerator(size, style_dim, n_mlp)
        sd = ref.state_dict()
        syn = synthetic_state_dict({k: v.shape for k, v in sd.items()}, seed=size)
        missing = ref.load_state_dict(syn, strict=False)
        assert not missing.unexpected_keys, missing
        ref = ref.to(dev).eval()
        z = torch.tensor(synthetic_z(batch, seed=size, dim_z=style_dim), dtype=torch.float32, device=dev)
        w = ref.style(z)
        # a W+ latent with per-layer variation so the latent index map is exercised
        gl = torch.Generator().manual_seed(100 + size)
        lat = (w[:, None, :] + 0.3 * torch.randn(batch, ref.n_latent, style_dim, generator=gl).to(dev)).detach()
        lat.requires_grad_(True)
        noise = [n.to(dev) for n in synthetic_noise(ref.num_layers, batch, seed=2)]
        img, _ = ref(lat, input_is_latent=True, noise=noise)
        probe = torch.randn(img.shape, generator=gl).to(dev)
        (glat,) = torch.autograd.grad((img * probe).sum(), lat)
        img_fixed, _ = ref(lat.detach(), input_is_latent=True, randomize_noise=False)
        tag = f"s{size}"
        gen_out[f"{tag}_cfg"] = np.array([size, style_dim, n_mlp, batch])
        gen_out[f"{tag}_w"] = w.detach().cpu().numpy()
        gen_out[f"{tag}_latent"] = lat.detach().cpu().numpy()
        gen_out[f"{tag}_image"] = img.detach().cpu().numpy()
        gen_out[f"{tag}_image_fixed_noise"] = img_fixed.detach().cpu().numpy()
        gen_out[f"{tag}_probe"] = probe.cpu().numpy()
        gen_out[f"{tag}_grad_latent"] = glat.cpu().numpy()
        print(f"generator size {size}: image range [{img.min().item():.3f}, {img.max().item():.3f}] std {img.std().item():.3f}",
              flush=True)
        del ref
    np.savez_compressed(os.path.join(out_dir, "ref_gpu_generator.npz"), **gen_out)

    # ---- walks --------------------------------------------------------------------------------
    _stub_missing_modules()
    walks = {}
    try:
        from graphs.stylegan_v2_real import transform_base as tb
        torch.manual_seed(5)
        n_latent, dim, batch, n_attr = 6, 64, 3, 2
        ws = [torch.randn(batch, dim, device=dev) for _ in range(n_latent)]
        alpha = torch.randn(batch, n_attr, device=dev)
        lin = tb.WalkLinearMultiW(dim, n_latent // 2 - 1, 1, ["a", "b"])
        with torch.no_grad():
            lin.w.copy_(synthetic_walk_w(n_attr, n_latent, dim, seed=3))
        lin = lin.to(dev)
        walks["ws"] = torch.stack(ws, 1).cpu().numpy()
        walks["alpha"] = alpha.cpu().numpy()
        walks["linear_w"] = lin.w.detach().cpu().numpy()
        walks["linear_out"] = torch.stack(lin(ws, alpha), 1).detach().cpu().numpy()
        walks["linear_out_layers"] = torch.stack(lin(ws, alpha, layers=[0, 3]), 1).detach().cpu().numpy()
        mlp = tb.WalkMlpMultiW(dim, n_latent // 2 - 1, 1, ["a"]).to(dev)
        for i, p in enumerate(mlp.parameters()):
            walks[f"mlp_p{i}"] = p.detach().cpu().numpy()
        walks["mlp_out"] = torch.stack(mlp(ws, alpha), 1).detach().cpu().numpy()
        nl = tb.WalkNonLinearW(dim, n_latent // 2 - 1, 1, ["a"]).to(dev)
        for i, p in enumerate(nl.parameters()):
            walks[f"nl_p{i}"] = p.detach().cpu().numpy()
        walks["nl_out"] = torch.stack(nl(ws, None, alpha, None), 1).detach().cpu().numpy()
        walks["nl_out_layers"] = torch.stack(nl(ws, None, alpha, None, layers=[1, 2]), 1).detach().cpu().numpy()
    except Exception as e:  # keep the op / generator fixtures even if the orchestration file cannot import
        print("walk fixtures skipped:", repr(e), flush=True)
    if walks:
        np.savez_compressed(os.path.join(out_dir, "ref_gpu_walks.npz"), **walks)
    print("done", flush=True)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
