"""Alpha schedules of the attribute transforms on the path (reference utils/transforms.py:634-735,
graphs/stylegan_v2_real/transform_op.py:65-77).  Training targets come from the *global* numpy RNG
exactly like the reference, so data-parallel ranks must seed it identically."""
import numpy as np


class _AttributeTransform:
    low, high = 0.0, 1.0

    def __init__(self, *args, **kwargs):
        self.alpha_max = 1

    def get_train_alpha(self, zs_batch, N_attr=40, trainEmbed=False):
        if trainEmbed:
            raise NotImplementedError("the embedding-bank walk is declared unused by the reference")
        alpha_val = np.random.uniform(self.low, self.high, N_attr)
        width = self._slider_width(N_attr)
        return np.ones((zs_batch.shape[0], width)) * alpha_val, alpha_val, None

    def scale_test_alpha_for_graph(self, alpha, zs_batch, **kwargs):
        return alpha * np.ones((zs_batch.shape[0], self.Nsliders))

    def test_alphas(self):
        return np.linspace(0, 1, 10)

    def vis_alphas(self, num_panels):
        return np.linspace(0, 1, num_panels)


class FaceTransform(_AttributeTransform):
    low, high = 0.0, 1.0

    def _slider_width(self, n_attr):
        return self.Nsliders  # broadcasts against alpha_val[N_attr]; Nsliders == 1 (transforms.py:659-663)

    def test_alphas(self):
        return np.linspace(0, 1, 9)


class SceneTransform(_AttributeTransform):
    low, high = -1.0, 1.0

    def _slider_width(self, n_attr):
        return n_attr
