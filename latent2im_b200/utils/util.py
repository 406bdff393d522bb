"""Reference utils/util.py:5-121 for the two transforms on the path (face, scene)."""
import os
from collections import OrderedDict

import numpy as np


def batch_input(graph_inputs, s):
    return {k: (v[s] if isinstance(v, np.ndarray) else v) for k, v in graph_inputs.items()}


def _read_attr_table(path):
    names, table = [], OrderedDict()
    with open(path) as f:
        for line in f:
            if line.strip():
                table[line.strip()] = len(names)
                names.append(line.strip())
    assert len(names) == 40, " len(attrList) should be 40"
    return names, table


def set_graph_kwargs(opt):
    kw = dict(lr=opt.learning_rate, walk_type=opt.walk_type, loss=opt.loss, trainEmbed=opt.trainEmbed)
    if opt.transform not in ("face", "scene"):
        raise NotImplementedError("transform %r is outside the accelerated path (face | scene)" % opt.transform)
    # the reference's default is '' (options/train_options.py:38) and the lists live in its dataset/ directory
    attr_path = opt.attrPath or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dataset",
                                             "attributes_celeba.txt" if opt.transform == "face" else "attributes_scene.txt")
    names, table = _read_attr_table(attr_path)
    kw["attrList"] = opt.attrList.split(",") if opt.attrList else names
    kw["attrTable"] = table
    try:
        kw["layers"] = opt.layers.split(",")
    except AttributeError:
        kw["layers"] = None
    if opt.walk_type.startswith("NN"):
        if getattr(opt.nn, "eps", None):
            kw["eps"] = opt.nn.eps
        if getattr(opt.nn, "num_steps", None):
            kw["N_f"] = opt.nn.num_steps
    if "stylegan" in opt.model:
        kw["stylegan_opts"] = opt.stylegan
    return kw
