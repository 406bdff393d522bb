"""``TrainOptions`` / ``VisOptions``: the reference's argparse + YAML front end (options/train_options.py:9-209,
options/vis_options.py:4-48) with the same flag names, the same nested ``opt.yml`` schema and the same output-directory
naming, on PyYAML (``oyaml`` is not needed: dicts keep insertion order) and working on Python >= 3.10, where argparse
titles the default group ``options`` instead of ``optional arguments`` (the reference then files every top-level flag
under ``opt.options.*``).  Added flags (all optional): ``--size``, ``--batch_size``, ``--dtype``,
``--epochs``, ``--max_iters``, ``--walk_mlp``."""
import argparse
import os
import sys
from collections import OrderedDict

import yaml

_TOPLEVEL_GROUPS = ("positional arguments", "optional arguments", "options")


class TrainOptions:
    def __init__(self):
        self.initialized = False
        self.parser = argparse.ArgumentParser("Training Parser")

    def initialize(self):
        p = self.parser
        p.add_argument("--config_file", type=argparse.FileType(mode="r"), help="configuration yml file")
        p.add_argument("--overwrite_config", action="store_true", help="overwrite config files if they exist")
        p.add_argument("--model", default="stylegan_v2_real", help="pretrained model to use")
        p.add_argument("--transform", default="face", help="transform operation: face | scene")
        p.add_argument("--num_samples", type=int, default=20000, help="number of latent z samples")
        p.add_argument("--loss", type=str, default="l2", choices=["l2", "lpips"])
        p.add_argument("--learning_rate", type=float, default=0.0001)
        p.add_argument("--walk_type", type=str, default="NNz", choices=["NNz", "linear"])
        p.add_argument("--models_dir", type=str, default="./models")
        p.add_argument("--model_save_freq", type=int, default=400)
        p.add_argument("--name", type=str)
        p.add_argument("--suffix", type=str)
        p.add_argument("--prefix", type=str)
        p.add_argument("--gpu", default="", type=str)
        p.add_argument("--trainEmbed", action="store_true")
        p.add_argument("--updateGAN", action="store_true")
        p.add_argument("--attrList", type=str)
        p.add_argument("--attrPath", type=str, default="")
        p.add_argument("--layers", type=str)
        p.add_argument("--no_content_loss", action="store_true")
        p.add_argument("--no_gan_loss", action="store_true")
        # additions of this implementation (the reference hard-codes 256 px / batch 4 / 10 epochs)
        p.add_argument("--size", type=int, default=None, help="generator resolution (default: constants.resolution)")
        p.add_argument("--batch_size", type=int, default=None, help="latents per step and GPU (default: constants.BATCH_SIZE)")
        p.add_argument("--dtype", type=str, default=None, choices=["bf16", "fp32"])
        p.add_argument("--epochs", type=int, default=10)
        p.add_argument("--max_iters", type=int, default=None, help="stop every epoch after this many iterations")
        p.add_argument("--walk_mlp", action="store_true", help="WalkMlpMultiW instead of WalkLinearMultiW (is_mlp, transform_base.py:291)")
        g = p.add_argument_group("nn", "parameters used to specify NN walk")
        g.add_argument("--eps", type=float)
        g.add_argument("--num_steps", type=int)
        g = p.add_argument_group("color", "parameters used for color walk")
        g.add_argument("--channel", type=int)
        g = p.add_argument_group("biggan", "parameters used for biggan walk")
        g.add_argument("--category", type=int)
        g = p.add_argument_group("stylegan", "parameters used for stylegan walk")
        g.add_argument("--dataset", default="scene")
        g.add_argument("--latent", default="w")
        g.add_argument("--truncation_psi", default=1.0)
        g = p.add_argument_group("pggan", "parameters used for pggan walk")
        g.add_argument("--dset", default="celebahq")
        self.initialized = True
        return p

    @staticmethod
    def _flatten(data):
        out = {}
        for k, v in data.items():
            if isinstance(v, dict):
                out.update(TrainOptions._flatten(v))
            else:
                out[k] = v
        return out

    def print_options(self, opt):
        d = OrderedDict()
        grouped = []
        for k, v in sorted(vars(opt).items()):
            if isinstance(v, argparse.Namespace):
                grouped.append((k, v))
            else:
                d[k] = v
        for k, v in grouped:
            d[k] = dict(sorted(vars(v).items()))
        out_dir = getattr(opt, "output_dir", "./")
        os.makedirs(out_dir, exist_ok=True)
        if not opt.overwrite_config:
            for f in ("opt.txt", "opt.yml"):
                assert not os.path.isfile(os.path.join(out_dir, f)), "config file exists, use --overwrite_config"
        with open(os.path.join(out_dir, "opt.txt"), "wt") as f:
            for k, v in d.items():
                f.write("{:>25}: {}\n".format(k, v))
        d["overwrite_config"] = False
        with open(os.path.join(out_dir, "opt.yml"), "wt") as f:
            yaml.safe_dump(dict(d), f, default_flow_style=False, sort_keys=False)

    def parse(self, argv=None, print_opt=True):
        if not self.initialized:
            self.initialize()
        argv = list(sys.argv[1:] if argv is None else argv)
        opt = self.parser.parse_args(argv)
        data = self._flatten(yaml.safe_load(opt.config_file)) if opt.config_file else {}
        option_strings = {o: a.dest for grp in self.parser._action_groups for a in grp._group_actions for o in a.option_strings}
        specified = {option_strings[x] for x in argv if x in option_strings}
        args = {}
        for grp in self.parser._action_groups:
            gd = {a.dest: data[a.dest] if (a.dest in data and a.dest not in specified) else getattr(opt, a.dest, None)
                  for a in grp._group_actions}
            if grp.title in _TOPLEVEL_GROUPS:
                args.update(gd)
            else:
                args[grp.title] = argparse.Namespace(**gd)
        args.pop("help", None)
        opt = argparse.Namespace(**args)
        delattr(opt, "config_file")
        if opt.name:
            out = opt.name
        else:
            out = "_".join([opt.model, opt.transform, opt.walk_type, "lr" + str(opt.learning_rate), opt.loss])
            if "stylegan" in opt.model:
                out += "_{}".format(opt.stylegan.latent)
        if opt.suffix:
            out += opt.suffix
        if opt.prefix:
            out = opt.prefix + out
        opt.output_dir = os.path.join(opt.models_dir, out)
        if print_opt:
            self.print_options(opt)
        self.opt = opt
        return opt


class VisOptions:
    def __init__(self):
        self.initialized = False
        self.parser = argparse.ArgumentParser("Visualization Parser")

    def initialize(self):
        p = self.parser
        p.add_argument("config_file", type=argparse.FileType(mode="r"), help="configuration yml file")
        p.add_argument("--save_path_w", type=str)
        p.add_argument("--save_path_gan", type=str)
        p.add_argument("--gpu", default="", type=str)
        p.add_argument("--noise_seed", type=int, default=0)
        p.add_argument("--output_dir")
        p.add_argument("--attrList", type=str)
        p.add_argument("--attrPath", type=str, default="")
        self.initialized = True
        return p

    def parse(self, argv=None):
        if not self.initialized:
            self.initialize()
        opt = self.parser.parse_args(argv)
        data = yaml.safe_load(opt.config_file)
        for k, v in list(data.items()):
            if isinstance(v, dict):
                data[k] = argparse.Namespace(**v)
        self.opt, self.data = opt, argparse.Namespace(**data)
        return self.opt, self.data
