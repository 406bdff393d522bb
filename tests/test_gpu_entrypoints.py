"""GPU: the train.py -> opt.yml + walk checkpoint -> vis_w.py round trip on a tiny configuration."""
import glob
import os

import pytest

pytestmark = pytest.mark.gpu


def test_train_then_vis_round_trip(tmp_path, monkeypatch):
    import train
    import vis_w
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    attr = os.path.join(root, "latent2im_b200", "dataset", "attributes_celeba.txt")
    monkeypatch.setenv("L2I_G_PATH", "/nonexistent")          # random-init generator / regressor
    monkeypatch.setenv("L2I_REG_PATH", "/nonexistent")
    out = train.main(["--model", "stylegan_v2_real", "--transform", "face", "--num_samples", "8", "--learning_rate", "1e-3",
                      "--latent", "w", "--walk_type", "linear", "--loss", "l2", "--attrList", "Smiling", "--attrPath", attr,
                      "--models_dir", str(tmp_path), "--overwrite_config", "--no_gan_loss", "--no_content_loss",
                      "--size", "32", "--batch_size", "2", "--dtype", "fp32", "--epochs", "1", "--max_iters", "2", "--log_every", "1"])
    ckpt = os.path.join(out, "model_w_1_final_walk_module.ckpt")
    assert os.path.exists(ckpt) and os.path.exists(os.path.join(out, "opt.yml"))
    log = open(os.path.join(out, "log.txt")).read()
    assert log.count("T, epc, bst, lss, alpha") == 2
    img_dir = vis_w.main([os.path.join(out, "opt.yml"), "--save_path_w", ckpt, "--noise_seed", "0", "--num_samples", "3",
                          "--num_panels", "4"])
    pngs = glob.glob(os.path.join(img_dir, "*.png"))
    assert len(pngs) == 3 and os.path.exists(os.path.join(img_dir, "index.html"))
    import PIL.Image
    im = PIL.Image.open(pngs[0])
    assert im.size == (4 * 33 + 1, 34)     # 4 panels of 32 px with 1 px padding
    out_amp = train.main(["--model", "stylegan_v2_real", "--transform", "face", "--num_samples", "4", "--learning_rate", "1e-3",
                          "--latent", "w", "--walk_type", "linear", "--loss", "l2", "--attrList", "Smiling", "--attrPath", attr,
                          "--models_dir", str(tmp_path / "amp"), "--overwrite_config", "--no_gan_loss", "--no_content_loss",
                          "--size", "32", "--batch_size", "2", "--dtype", "bf16", "--epochs", "1", "--max_iters", "1", "--amp"])
    assert os.path.exists(os.path.join(out_amp, "model_w_1_final_walk_module.ckpt"))
    img_dir2 = vis_w.main([os.path.join(out, "opt.yml"), "--save_path_w", ckpt, "--noise_seed", "0", "--num_samples", "2",
                           "--num_panels", "3", "--cache_original", "--output_dir", str(tmp_path / "cached")])
    assert len(glob.glob(os.path.join(img_dir2, "*.png"))) == 2


def test_train_multi_attr_two_attributes(tmp_path, monkeypatch):
    """train_multi_attr.py (BASELINE config 4 caller): two attributes, clamped delta targets, loss_values.npy."""
    import numpy as np

    import train_multi_attr
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    attr = os.path.join(root, "latent2im_b200", "dataset", "attributes_celeba.txt")
    monkeypatch.setenv("L2I_G_PATH", "/nonexistent")
    monkeypatch.setenv("L2I_REG_PATH", "/nonexistent")
    out = train_multi_attr.main(["--model", "stylegan_v2_real", "--transform", "scene", "--num_samples", "8", "--learning_rate", "1e-3",
                                 "--latent", "w", "--walk_type", "linear", "--loss", "l2", "--attrList", "Smiling,Male", "--attrPath", attr,
                                 "--models_dir", str(tmp_path), "--overwrite_config", "--no_gan_loss", "--no_content_loss",
                                 "--size", "32", "--batch_size", "2", "--dtype", "fp32", "--epochs", "1", "--max_iters", "2", "--log_every", "1"])
    assert os.path.exists(os.path.join(out, "model_w_1_final_walk_module.ckpt"))
    losses = np.load(os.path.join(out, "loss_values.npy"))
    assert losses.shape == (2,) and np.isfinite(losses).all()
    assert open(os.path.join(out, "log.txt")).read().count("alpha night:") == 2
