"""Bring-up aid: compares the tcgen05 conv path with the CUDA-core path and the CPU oracle, size by size."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
from latent2im_b200.synthetic import load_synthetic, synthetic_noise
from oracle import GeneratorSpec, generator_forward_ref


def psnr(a, b):
    return 10 * torch.log10(4.0 / ((a - b) ** 2).mean()).item()


def run(size, dim, batch, impl):
    os.environ["L2I_CONV_IMPL"] = impl
    gen = load_synthetic(Generator(size, dim, 1), seed=0).cuda()
    gen.set_native(dtype=torch.bfloat16)
    spec = GeneratorSpec(size=size, style_dim=dim, n_mlp=1)
    lat = torch.randn(batch, spec.n_latent, dim, generator=torch.Generator().manual_seed(1))
    noise = synthetic_noise(spec.num_layers, batch)
    img, _ = gen(lat.cuda(), input_is_latent=True, noise=[n.cuda() for n in noise])
    torch.cuda.synchronize()
    skips = {}
    k = spec.log_size - 2
    for kk in (k - 1, k):
        if kk >= 0:
            skips[kk] = gen.read_activation(f"skip.{kk}").cpu().double()
    sd = {kk: v.double().cpu() for kk, v in gen.state_dict().items()}
    return img.cpu().double(), skips, (sd, lat, noise, spec)


if __name__ == "__main__":
    sizes = [int(s) for s in sys.argv[1].split(",")] if len(sys.argv) > 1 else [8, 16, 32, 64]
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    for size in sizes:
        a, sa, (sd, lat, noise, spec) = run(size, 64, batch, "simt")
        b, sb, _ = run(size, 64, batch, "tc")
        ref, inter = generator_forward_ref(sd, lat.double(), noise, spec, return_intermediates=True)
        print(f"size {size:4d} batch {batch}: simt-vs-oracle {psnr(a, ref):6.1f} dB  tc-vs-oracle {psnr(b, ref):6.1f} dB  "
              f"tc-vs-simt {psnr(a, b):6.1f} dB  max|tc-simt| {(a - b).abs().max().item():.3e}", flush=True)
        for kk in sb:
            name = "to_rgb1" if kk == 0 else f"to_rgbs.{kk - 1}"
            print(f"     skip.{kk}: tc-vs-oracle max {(sb[kk] - inter[name]).abs().max().item():.3e}  "
                  f"simt-vs-oracle max {(sa[kk] - inter[name]).abs().max().item():.3e}", flush=True)
