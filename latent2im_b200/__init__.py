"""latent2im_b200 - B200-native (sm_100a) implementation of the Latent2im StyleGAN2 latent-walk hot path.

The package mirrors the reference's import layout under ``latent2im_b200.graphs`` and can alias
it to the reference's top-level ``graphs`` package name (``install_dropin()``), so walk-module
pickles written by the reference (``torch.save(self.walk, ...)``, transform_base.py:499) resolve.
"""
import importlib
import sys

__version__ = "0.1.0"


def install_dropin():
    """Registers ``graphs`` / ``graphs.stylegan_v2_real.*`` as aliases of this package's modules."""
    names = ["graphs", "graphs.stylegan_v2_real", "graphs.stylegan_v2_real.op",
             "graphs.stylegan_v2_real.op.fused_act", "graphs.stylegan_v2_real.op.upfirdn2d",
             "graphs.stylegan_v2_real.networks", "graphs.stylegan_v2_real.transform_base",
             "graphs.stylegan_v2_real.graph_util", "graphs.stylegan_v2_real.constants",
             "graphs.stylegan_v2_real.transform_op", "graphs.transform_graph_scene"]
    for n in names:
        try:
            sys.modules[n] = importlib.import_module("latent2im_b200." + n)
        except ModuleNotFoundError:
            pass
