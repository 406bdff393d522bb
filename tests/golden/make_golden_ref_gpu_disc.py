"""Golden fixture of the UNMODIFIED reference Discriminator on a CUDA device.

    python oracle/stage_reference.py                   # in the build container, once (+ prebuilds the ops when imported there)
    gpurun -- python tests/golden/make_golden_ref_gpu_disc.py gpurun_out/golden
    cp gpurun_out/golden/ref_gpu_discriminator.npz tests/golden/

Weights and images are rebuilt from seeds (``latent2im_b200.synthetic``), so only logits, the cotangent and the
input gradient are stored.  TF32 is disabled: the fixture is the reference's fp32 arithmetic.
"""
import json
import os
import sys

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

CASES = [
    {"name": "d32", "size": 32, "cm": 2, "batch": 4, "seed": 0, "grad": True},
    {"name": "d64", "size": 64, "cm": 1, "batch": 8, "seed": 1, "grad": False},   # two minibatch-stddev groups
]


def image(batch, size, seed):
    g = torch.Generator().manual_seed(seed)
    return 0.5 * torch.randn(batch, 3, size, size, generator=g, dtype=torch.float32)


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    from graphs.stylegan_v2_real.networks import Discriminator as RefDiscriminator
    from latent2im_b200.synthetic import synthetic_discriminator_state_dict
    dev = torch.device("cuda")
    out = {"cases": json.dumps(CASES)}
    for c in CASES:
        d = RefDiscriminator(c["size"], channel_multiplier=c["cm"])
        sd = synthetic_discriminator_state_dict({k: v.shape for k, v in d.state_dict().items()}, c["seed"])
        missing = d.load_state_dict(sd, strict=False)
        assert all(k.endswith(".kernel") for k in missing.missing_keys) and not missing.unexpected_keys, missing
        d = d.to(dev).eval()
        x = image(c["batch"], c["size"], 100 + c["seed"]).to(dev).requires_grad_(True)
        logits = d(x)
        out[c["name"] + "_logits"] = logits.detach().cpu().numpy()
        if c["grad"]:
            g = torch.Generator().manual_seed(7)
            coef = torch.randn(c["batch"], 1, generator=g, dtype=torch.float32)
            (logits * coef.to(dev)).sum().backward()
            out[c["name"] + "_coef"] = coef.numpy()
            out[c["name"] + "_grad"] = x.grad.detach().cpu().numpy()
        print(c["name"], "logits", logits.detach().flatten().tolist()[:4], flush=True)
    np.savez_compressed(os.path.join(out_dir, "ref_gpu_discriminator.npz"), **out)
    print("wrote", os.path.join(out_dir, "ref_gpu_discriminator.npz"))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
