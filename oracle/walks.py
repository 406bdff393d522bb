"""Oracle restatement of the latent-walk modules (test infrastructure only).

Reference lines followed (graphs/stylegan_v2_real/):
  z sampling ................. graph_util.py:5-8
  WalkLinearMultiW.forward ... transform_base.py:151-165
  WalkMlpMultiW.forward ...... transform_base.py:181-204 (the ``layers`` branch calls
                               ``self.linear(input[i], 1)``, a TypeError as shipped; the oracle
                               implements the evident intent: same MLP, only on the chosen layers)
  WalkNonLinearW.forward ..... transform_base.py:219-243
  get_alphas ................. transform_base.py:405-408
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def z_sample_ref(batch_size: int, seed: int = 0, dim_z: int = 512) -> np.ndarray:
    return np.random.RandomState(seed).randn(batch_size, dim_z)


def get_alphas_ref(alpha_org, alpha_target):
    return alpha_target - alpha_org


def walk_linear_ref(ws, alpha, w_param, layers=None):
    """``ws``: list of n_latent tensors [B, D]; ``alpha``: [B, A]; ``w_param``: [A, n_latent, D]."""
    out = []
    for i, w_i in enumerate(ws):
        if layers is None or i in layers:
            out.append(w_i + alpha.to(w_i.dtype) @ w_param[:, i, :].to(w_i.dtype))
        else:
            out.append(w_i)
    return out


def _mlp(x, params, slope=0.2):
    """nn.Sequential(Linear, LeakyReLU(0.2), ..., Linear); ``params`` = [(W, b), ...]."""
    n = len(params)
    for j, (w, b) in enumerate(params):
        x = F.linear(x, w.to(x.dtype), b.to(x.dtype))
        if j + 1 < n:
            x = F.leaky_relu(x, slope)
    return x


def walk_mlp_ref(ws, alpha, mlp_params, layers=None):
    """``out_i = in_i + alpha[:, 0:1] * MLP(in_i)``; MLP = 512->1024->1024->512."""
    al = alpha[:, 0:1]
    out = []
    for i, w_i in enumerate(ws):
        if layers is None or i in layers:
            out.append(w_i + al.to(w_i.dtype) * _mlp(w_i, mlp_params))
        else:
            out.append(w_i)
    return out


def walk_nonlinear_ref(ws, alpha, embed_params, mlp_params, layers=None):
    """``e = embed(alpha[:, 0:1].repeat(1, 10))``; ``d = MLP(cat[e, in_i])``;
    ``out_i = in_i + d / ||d||`` (no normalisation when ``layers`` is given, :237-239)."""
    al = alpha[:, 0:1]
    ew, eb = embed_params
    out = []
    for i, w_i in enumerate(ws):
        e = F.linear(al.to(w_i.dtype).repeat(1, 10), ew.to(w_i.dtype), eb.to(w_i.dtype))
        if layers is None:
            d = _mlp(torch.cat([e, w_i], 1), mlp_params)
            out.append(w_i + d / torch.norm(d, dim=1, keepdim=True))
        elif i in layers:
            out.append(w_i + _mlp(torch.cat([e, w_i], 1), mlp_params))
        else:
            out.append(w_i)
    return out
