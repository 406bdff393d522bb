#!/usr/bin/env python
"""Panel rendering entry point - drop-in for the reference's ``vis_w.py`` (vis_w.py:21-118): loads ``opt.yml`` and a
walk checkpoint, sweeps ``linspace(min_alpha, max_alpha, num_panels)`` targets and writes one PNG grid per sample plus
``index.html``.

    python vis_w.py models_celeba/<run>/opt.yml --noise_seed 0 --num_samples 30 --num_panels 10 \\
        --save_path_w models_celeba/<run>/model_w_10_final_walk_module.ckpt [--size 1024 --batch_size 32]
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main(argv=None):
    import torch
    from latent2im_b200 import graphs
    from latent2im_b200.options import VisOptions
    from latent2im_b200.utils import html, util

    v = VisOptions()
    v.initialize()
    v.parser.add_argument("--num_samples", type=int, default=10)
    v.parser.add_argument("--num_panels", type=int, default=7)
    v.parser.add_argument("--max_alpha", type=float, default=1)
    v.parser.add_argument("--min_alpha", type=float, default=0)
    v.parser.add_argument("--layers", type=str, default=None)
    v.parser.add_argument("--trainEmbed", action="store_true")
    v.parser.add_argument("--updateGAN", action="store_true")
    v.parser.add_argument("--size", type=int, default=None)
    v.parser.add_argument("--batch_size", type=int, default=None)
    v.parser.add_argument("--allow_random_init", action="store_true", help="run on random-init G / R when no checkpoint is found")
    v.parser.add_argument("--cache_original", action="store_true",
                          help="compute G(w) and R(G(w)) once per batch instead of once per panel (the reference recomputes them)")
    opt, conf = v.parse(argv)
    if opt.gpu:
        os.environ["CUDA_VISIBLE_DEVICES"] = opt.gpu       # before the first CUDA call: the visible set is latched by it
    assert torch.cuda.is_available(), "vis_w.py needs a CUDA device (there is no CPU fallback)"
    out_dir = opt.output_dir or os.path.join(conf.output_dir, "images")
    os.makedirs(out_dir, exist_ok=True)
    constants = importlib.import_module("latent2im_b200.graphs." + conf.model + ".constants")
    graph_util = importlib.import_module("latent2im_b200.graphs." + conf.model + ".graph_util")
    size = opt.size or getattr(conf, "size", None)
    if size:
        constants.resolution = size
    bs = opt.batch_size or getattr(conf, "batch_size", None)
    if bs:
        constants.BATCH_SIZE = bs
    if getattr(conf, "dtype", None):
        constants.compute_dtype = conf.dtype
    constants.walk_is_mlp = bool(getattr(conf, "walk_mlp", False))
    if opt.allow_random_init:
        constants.allow_random_init = True
    g = graphs.find_model_using_name(conf.model, conf.transform)(**util.set_graph_kwargs(conf))
    g.load_multi_models(opt.save_path_w, None, trainEmbed=opt.trainEmbed, updateGAN=opt.updateGAN)
    inputs = graph_util.graph_input(g, opt.num_samples, seed=opt.noise_seed)
    epochs = os.path.basename(opt.save_path_w).split("_")[2]
    name = conf.attrList.strip().split(",")[0]
    layers = None if opt.layers in (None, "None") else [int(i) for i in opt.layers.split(",")]
    filename = os.path.join(out_dir, "w_{}_seed{}_{}_max{}_min{}".format(epochs, opt.noise_seed, name, opt.max_alpha, opt.min_alpha))
    b = constants.BATCH_SIZE
    for start in range(0, opt.num_samples, b):
        s = slice(start, min(opt.num_samples, start + b))
        batch = util.batch_input(inputs, s)
        to_graph, to_target = g.vis_image_batch(batch, filename, s.start, num_panels=opt.num_panels, max_alpha=opt.max_alpha,
                                                min_alpha=opt.min_alpha, wgt=True)
        g.vis_multi_image_batch_alphas(batch, filename, alphas_to_graph=to_graph, alphas_to_target=to_target, layers=layers,
                                       batch_start=s.start, name=name, wgt=False, wmask=False, trainEmbed=opt.trainEmbed,
                                       computeL2=False, given_w=None, cache_original=opt.cache_original)
    html.make_html(out_dir)
    return out_dir


if __name__ == "__main__":
    main()
