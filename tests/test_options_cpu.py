"""CPU: the opt.yml contract between train.py (TrainOptions) and vis_w.py (VisOptions) - reference
options/train_options.py:118-121,150-202 and options/vis_options.py:30-48."""
import os

import yaml

from latent2im_b200.options import TrainOptions, VisOptions


def test_train_options_write_nested_yaml_and_vis_reads_it(tmp_path):
    t = TrainOptions()
    opt = t.parse(["--model", "stylegan_v2_real", "--transform", "face", "--walk_type", "linear", "--latent", "w",
                   "--learning_rate", "1e-4", "--attrList", "Smiling", "--attrPath", "x.txt", "--models_dir", str(tmp_path),
                   "--overwrite_config", "--no_gan_loss", "--no_content_loss", "--size", "1024", "--batch_size", "16"])
    # output dir naming: <models_dir>/<model>_<transform>_<walk_type>_lr<lr>_<loss>_<latent>  (train_options.py:183-202)
    assert opt.output_dir == os.path.join(str(tmp_path), "stylegan_v2_real_face_linear_lr0.0001_l2_w")
    assert opt.stylegan.latent == "w" and opt.no_gan_loss and opt.size == 1024      # top-level flags stay top-level (Python >= 3.10 too)
    y = yaml.safe_load(open(os.path.join(opt.output_dir, "opt.yml")))
    assert y["walk_type"] == "linear" and y["stylegan"]["latent"] == "w" and y["nn"] == {"eps": None, "num_steps": None}
    assert y["overwrite_config"] is False
    v = VisOptions()
    vopt, conf = v.parse([os.path.join(opt.output_dir, "opt.yml"), "--save_path_w", "model_w_10_final_walk_module.ckpt", "--noise_seed", "3"])
    assert conf.model == "stylegan_v2_real" and conf.stylegan.latent == "w" and conf.attrList == "Smiling" and vopt.noise_seed == 3
    # config file + command line: explicit flags win, everything else comes from the file (train_options.py:150-169)
    t2 = TrainOptions()
    opt2 = t2.parse(["--config_file", os.path.join(opt.output_dir, "opt.yml"), "--learning_rate", "0.01", "--overwrite_config"])
    assert opt2.learning_rate == 0.01 and opt2.walk_type == "linear" and opt2.stylegan.latent == "w"


def test_attribute_table_lookup():
    from latent2im_b200.utils import util
    import argparse
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "latent2im_b200", "dataset", "attributes_celeba.txt")
    opt = argparse.Namespace(learning_rate=1e-4, walk_type="linear", loss="l2", trainEmbed=False, transform="face", attrPath=path,
                             attrList="Smiling", layers=None, model="stylegan_v2_real", stylegan=argparse.Namespace(latent="w"),
                             nn=argparse.Namespace(eps=None, num_steps=None))
    kw = util.set_graph_kwargs(opt)
    assert kw["attrList"] == ["Smiling"] and kw["attrTable"]["Smiling"] == 31 and len(kw["attrTable"]) == 40
