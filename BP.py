#!/usr/bin/env python
"""GAN inversion entry point - drop-in for the reference's ``BP.py`` (same flags and artefacts) over the B200-native
generator forward / data-gradient kernels (``latent2im_b200.inversion``).

    python BP.py --batch_size 1 --optimizer Adam --dataset ffhq --n_loops 4000 --path ./data/face --save_path ./results_face

For every batch of the image folder ``--path`` (class sub-directories, torchvision ``ImageFolder`` layout; resized and
centre-cropped to ``--resolution``, mapped to [-1, 1]) it optimises ``w`` in W+ from ``mean_latent(4096)`` and writes
``org_<i>.png``, ``<i>_final.png``, ``latent/<i>_w.npy`` and ``loss_back.npy`` (BP.py:119-171, 187-203, 236-260).
The generator checkpoint comes from ``L2I_G_PATH`` / ``--ckpt_path`` (the reference hard-codes two private paths,
BP.py:219-227); without one the random-init generator is used.  ``--vgg_path`` (torchvision ``vgg16`` state_dict) enables
the Gram-matrix perceptual term, which otherwise needs the network for ``pretrained=True``.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch


def build_parser():
    p = argparse.ArgumentParser(description="Backprop")
    p.add_argument("--latent_dim", type=int, default=512)
    p.add_argument("--batch_size", type=int, default=1)
    p.add_argument("--ckpt_path", type=str, default=os.environ.get("L2I_G_PATH", ""))
    p.add_argument("--allow_random_init", action="store_true")
    p.add_argument("--gpu", type=str, default="0")
    p.add_argument("--n_loops", type=int, default=500)
    p.add_argument("--resolution", type=int, default=256, choices=[32, 64, 128, 256, 512, 1024])
    p.add_argument("--optimizer", "--optim", dest="optimizer", type=str, choices=["Adam", "GD"], default="Adam")
    p.add_argument("--dataset", type=str, choices=["ffhq", "scene", "anime"], default="ffhq")
    p.add_argument("--path", type=str, required=True)
    p.add_argument("--save_path", type=str, default="./results")
    p.add_argument("--lr", type=float, default=0.01)
    p.add_argument("--dtype", type=str, choices=["fp32", "bf16"], default="fp32")
    p.add_argument("--vgg_path", type=str, default="")
    p.add_argument("--max_batches", type=int, default=None)
    return p


def load_image_folder(path, resolution):
    """[(tensor [3, R, R] in [-1, 1], class index)] in ``ImageFolder`` order: Resize(R) + CenterCrop(R) + ToTensor +
    Normalize(0.5, 0.5) (BP.py:237-244)."""
    import torchvision
    from torchvision import transforms
    tf = transforms.Compose([transforms.Resize(resolution), transforms.CenterCrop(resolution), transforms.ToTensor(),
                             transforms.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))])
    return torchvision.datasets.ImageFolder(path, transform=tf)


def save_column(images, path):
    """``utils.save_image(result, nrow=1, normalize=True, range=(-1, 1))`` (BP.py:205-216): the batch stacked vertically."""
    from PIL import Image
    x = ((images.detach().float().cpu().clamp(-1, 1) + 1) / 2 * 255).round().to(torch.uint8)      # [N, 3, H, W]
    col = x.permute(0, 2, 3, 1).reshape(-1, x.shape[3], 3).numpy()
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    Image.fromarray(col).save(path)


def run(args, generator=None, device=None):
    from latent2im_b200.inversion import GramPerceptualLoss, invert
    if generator is None:
        from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
        assert torch.cuda.is_available(), "BP.py needs a CUDA device (there is no CPU fallback)"
        device = torch.device("cuda", torch.cuda.current_device())
        generator = Generator(args.resolution, args.latent_dim, 8)
        if args.ckpt_path and os.path.exists(args.ckpt_path):
            res = generator.load_state_dict(torch.load(args.ckpt_path, map_location="cpu", weights_only=False)["g_ema"], strict=False)
            bad = [k for k in res.missing_keys if not (k.endswith(".kernel") or k.startswith("noises."))] + list(res.unexpected_keys)
            if bad:
                print(f"BP.py: WARNING: {args.ckpt_path}: missing / unexpected keys {bad[:8]}", file=sys.stderr)
        elif os.environ.get("L2I_ALLOW_RANDOM_INIT", "0") in ("0", "") and not getattr(args, "allow_random_init", False):
            raise FileNotFoundError(f"generator checkpoint not found: {args.ckpt_path!r} (--allow_random_init to invert on random weights)")
        else:
            print("BP.py: WARNING: no generator checkpoint: inverting against RANDOM-INIT weights", file=sys.stderr)
        generator.set_native(dtype={"fp32": torch.float32, "bf16": torch.bfloat16}[args.dtype], max_batch=args.batch_size)
        generator = generator.to(device).eval()
    perceptual = GramPerceptualLoss(args.vgg_path).to(device) if args.vgg_path else None
    data = load_image_folder(args.path, args.resolution)
    loader = torch.utils.data.DataLoader(data, batch_size=args.batch_size)
    os.makedirs(args.save_path, exist_ok=True)
    with torch.no_grad():
        mean_latent = generator.mean_latent(4096)
    all_losses = []
    for i, (batch, _) in enumerate(loader):
        if args.max_batches is not None and i >= args.max_batches:
            break
        save_column(batch, os.path.join(args.save_path, "org_%d.png" % i))
        target = batch.to(device)
        w, losses = invert(generator, target, n_loops=args.n_loops, lr=args.lr, optim=args.optimizer, mean_latent=mean_latent,
                           extra_loss=(lambda out, tgt: perceptual(tgt, out)) if perceptual is not None else None)
        with torch.no_grad():
            out, _ = generator(w, input_is_latent=True)
        save_column(out, os.path.join(args.save_path, "%d_final.png" % i))
        os.makedirs(os.path.join(args.save_path, "latent"), exist_ok=True)
        np.save(os.path.join(args.save_path, "latent", "%d_w.npy" % i), w.cpu().numpy())
        all_losses.extend(losses.double().cpu().tolist())
        np.save(os.path.join(args.save_path, "loss_back.npy"), np.array(all_losses))          # cumulative, as BP.py:190
        print("[%d] %d loops: loss %.4f -> %.4f" % (i, args.n_loops, float(losses[0]), float(losses[-1])))
    return args.save_path


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.gpu and "CUDA_VISIBLE_DEVICES" not in os.environ:
        os.environ["CUDA_VISIBLE_DEVICES"] = args.gpu
    return run(args)


if __name__ == "__main__":
    main()
