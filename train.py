#!/usr/bin/env python
"""Walk-module training entry point - drop-in for the reference's ``train.py`` (same flags, same ``opt.yml`` /
``log.txt`` / ``model_w_<epoch>_walk_module.ckpt`` artefacts) over the B200-native hot path.

    python train.py --model stylegan_v2_real --transform face --num_samples 20000 --learning_rate 1e-4 --latent w \\
        --walk_type linear --loss l2 --attrList Smiling --attrPath latent2im_b200/dataset/attributes_celeba.txt \\
        --models_dir ./models_celeba --overwrite_config --no_gan_loss --no_content_loss [--size 1024 --batch_size 16]
    torchrun --nproc-per-node 8 train.py ...            # data parallel: latents sharded, walk gradient all-reduced

Loop structure follows reference train.py:25-132.  Differences: G-forward #1 / R-forward #1 run without autograd
(zero contribution to the walk gradient); the loss is synchronised to the host only every
``--log_every`` iterations instead of every iteration (train.py:110); TensorBoard is optional.
"""
import logging
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


def main(argv=None, multi_attr=False):
    from latent2im_b200 import graphs, parallel
    from latent2im_b200.options import TrainOptions
    from latent2im_b200.utils import image, util

    t = TrainOptions()
    t.initialize()
    t.parser.add_argument("--log_every", type=int, default=10)
    t.parser.add_argument("--allow_random_init", action="store_true", help="run on random-init G / R when no checkpoint is found")
    t.parser.add_argument("--amp", action="store_true", help="run the stock ResNet-50 regressor under bf16 autocast + channels_last")
    rank, world, local = parallel.world_info()
    opt = t.parse(argv, print_opt=(rank == 0))
    if multi_attr and not any(a.startswith("--epochs") for a in (sys.argv[1:] if argv is None else argv)):
        opt.epochs = 3                                        # train_multi_attr.py:54 hard-codes three epochs
    if opt.gpu and world == 1:
        os.environ["CUDA_VISIBLE_DEVICES"] = opt.gpu       # before the first CUDA call: the visible set is latched by it
    assert torch.cuda.is_available(), "train.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import importlib
    constants = importlib.import_module("latent2im_b200.graphs." + opt.model + ".constants")
    graph_util = importlib.import_module("latent2im_b200.graphs." + opt.model + ".graph_util")
    if opt.size:
        constants.resolution = opt.size
    if opt.batch_size:
        constants.BATCH_SIZE = opt.batch_size
    if opt.dtype:
        constants.compute_dtype = opt.dtype
    constants.walk_is_mlp = bool(opt.walk_mlp)
    if opt.allow_random_init:
        constants.allow_random_init = True
    if opt.amp:
        constants.reg_amp = True

    kw = util.set_graph_kwargs(opt)
    g = graphs.find_model_using_name(opt.model, opt.transform)(**kw)
    out_dir = train_loop(g, opt, constants, graph_util, kw["attrList"], rank, world, multi_attr=multi_attr)
    if world > 1:
        dist.destroy_process_group()
    return out_dir


def multi_attr_targets(alpha_org, alpha_delta):
    """Intended semantics of ``train_multi_attr.py`` (reference train_multi_attr.py:113 unpacks two values; the
    two-value ``get_alphas`` is graphs/pggan/transform_base.py:358-364): the sampled value is a DELTA on the current
    regressor prediction, the regression target is the prediction moved by it and clamped to [0, 1], and the walk is
    driven by the clamped difference."""
    alpha_target = torch.clamp(alpha_org + alpha_delta, min=0, max=1)
    return alpha_target, alpha_target - alpha_org


def train_loop(g, opt, constants, graph_util, attr_list, rank=0, world=1, multi_attr=False):
    """The iteration loop of reference train.py:25-132 / train_multi_attr.py:46-170 over any object with the
    TransformGraph method set (``g.device`` decides where the batches go)."""
    from latent2im_b200.utils import image
    out_dir = opt.output_dir
    log = logging.getLogger("latent2im.train")
    if rank == 0:
        os.makedirs(out_dir, exist_ok=True)
        log.setLevel(logging.INFO)
        for h in list(log.handlers):
            log.removeHandler(h)
        log.addHandler(logging.FileHandler(os.path.join(out_dir, "log.txt"), mode="w"))
    b = constants.BATCH_SIZE
    global_b = b * world
    losses = []
    tag = "alpha night" if multi_attr else "alpha"          # the two reference scripts label the column differently
    for epoch in range(opt.epochs):
        inputs = graph_util.graph_input(g, opt.num_samples, seed=epoch)      # identical on every rank (seeded numpy)
        np.random.seed(100003 * epoch + 17)                                   # same target stream on every rank
        iters = opt.num_samples // global_b
        if opt.max_iters:
            iters = min(iters, opt.max_iters)
        for i in range(iters):
            t0 = time.time()
            rows = slice(i * global_b + rank * b, i * global_b + (rank + 1) * b)
            zs = inputs["z"][rows]
            z = torch.Tensor(zs).to(g.device)
            with torch.no_grad():
                w = g.get_w(z)
                out_zs = g.get_logits({"w": w})
                alpha_org = g.get_reg_preds(out_zs)
            ag, at, _ = g.get_train_alpha(zs, N_attr=len(attr_list), trainEmbed=opt.trainEmbed)
            ag_t = torch.tensor(ag).float().to(g.device)
            if multi_attr:
                target, eps = multi_attr_targets(alpha_org, ag_t)
            else:
                target, eps = ag_t, g.get_alphas(alpha_org, ag_t)
            w_new = g.get_w_new_tensor(w, eps, layers=opt.layers)
            out = g.get_logits({"w": w_new})
            loss = g.optimizeParametersAll({"w": w_new, "org": out_zs, "logit": out, "alpha": target}, trainEmbed=opt.trainEmbed,
                                           updateGAN=opt.updateGAN, no_content_loss=opt.no_content_loss, no_gan_loss=opt.no_gan_loss)
            if multi_attr:
                losses.append(loss.detach())     # kept on the device: loss_values.npy is written once, at the end
            if rank == 0 and i % opt.log_every == 0:
                log.info("T, epc, bst, lss, {}: {}, {}, {}, {}, {}".format(tag, time.time() - t0, epoch, i * global_b,
                                                                          loss.item(), round(float(at[0]), 2)))
            if rank == 0 and i % opt.model_save_freq == 0:
                u8 = g.clip_ims(out.detach().float().cpu().numpy()).transpose(0, 2, 3, 1)
                image.save_im(image.imgrid(u8, cols=len(u8)), os.path.join(out_dir, "results", "{}_{}_logit_{:.2f}".format(epoch, i * global_b, float(at[0]))))
        if rank == 0:
            g.save_multi_models("{}/model_w_{}".format(out_dir, epoch), None)
    if rank == 0:
        g.save_multi_models("{}/model_w_{}_final".format(out_dir, opt.epochs), None)
        if multi_attr:   # train_multi_attr.py:224-226 (the matplotlib plot of the same values is left to the user)
            np.save(os.path.join(out_dir, "loss_values.npy"),
                    torch.stack(losses).double().cpu().numpy() if losses else np.zeros(0))
        for h in list(log.handlers):
            h.close()
    return out_dir


if __name__ == "__main__":
    main()
