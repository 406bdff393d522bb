"""``faceGraph`` / ``SceneGraph`` (reference graphs/transform_graph_scene.py:5-125): the orchestrator
combined with an attribute alpha schedule."""
import importlib

import numpy as np


def get_transform_graphs(model):
    pkg = __name__.rsplit(".", 1)[0] + "." + model
    base = importlib.import_module(pkg + ".transform_base")
    op = importlib.import_module(pkg + ".transform_op")
    constants = importlib.import_module(pkg + ".constants")

    def make(name, op_cls):
        class Graph(base.PixelTransform, op_cls):
            def __init__(self, lr=0.001, walk_type="NNz", loss="l2", eps=1.41, N_f=4, **kwargs):
                self.walk_type = walk_type
                self.num_channels = constants.NUM_CHANNELS
                self.Nsliders = 1
                self.img_size = constants.resolution
                base.PixelTransform.__init__(self, lr, walk_type, 1, loss, eps, N_f, **kwargs)
                op_cls.__init__(self)

            def vis_image_batch(self, graph_inputs, filename, batch_start, wgt=False, wmask=False, num_panels=7,
                                max_alpha=None, min_alpha=None, N_attr=40):
                zs = graph_inputs["z"]
                lo, hi = (min_alpha, max_alpha) if (max_alpha is not None and min_alpha is not None) else (0, 1)
                alphas = np.linspace(lo, hi, num_panels)
                return [self.scale_test_alpha_for_graph(a, zs) for a in alphas], list(alphas)

        Graph.__name__ = Graph.__qualname__ = name
        return Graph

    return [make("SceneGraph", op.SceneTransform), make("faceGraph", op.FaceTransform)]
