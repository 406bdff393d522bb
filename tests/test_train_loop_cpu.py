"""Host logic of train.py / train_multi_attr.py on CPU: the iteration loop driven through a stand-in graph object
(the TransformGraph method set over tiny CPU tensors), so targets, clamping, sharding, logging and artefacts are
checked without a GPU."""
import os
import types

import numpy as np
import torch

import train


class _StubGraph:
    """alpha_org = sigmoid(mean of w) per attribute; the 'walk' adds a learnable offset times eps."""

    def __init__(self, n_attr):
        self.device = torch.device("cpu")
        self.n_attr = n_attr
        self.offset = torch.nn.Parameter(torch.zeros(n_attr, 8))
        self.opt = torch.optim.SGD([self.offset], lr=0.5)
        self.calls = []
        self.saved = []

    def get_w(self, z):
        return [z[:, :8]] * 4

    def get_logits(self, d):
        w = d["w"]
        w = torch.stack(list(w), 1) if isinstance(w, (list, tuple)) else w
        return w.mean(1)[:, :, None, None].expand(-1, -1, 2, 2)[:, :3]          # [B, 3, 2, 2] "image"

    def get_reg_preds(self, img):
        m = img.mean((1, 2, 3))
        return torch.sigmoid(torch.stack([m * (k + 1) for k in range(self.n_attr)], 1))

    def get_train_alpha(self, zs, N_attr=1, trainEmbed=False):
        a = np.random.uniform(-1, 1, N_attr)
        return np.ones((zs.shape[0], N_attr)) * a, a, None

    def get_alphas(self, alpha_org, target):
        return target - alpha_org

    def get_w_new_tensor(self, ws, eps, layers=None):
        return [w + eps @ self.offset for w in ws]

    def optimizeParametersAll(self, feed, trainEmbed, updateGAN, no_content_loss=False, no_gan_loss=False):
        self.calls.append({k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in feed.items() if k in ("alpha", "org")})
        self.opt.zero_grad()
        loss = ((self.get_reg_preds(feed["logit"]) - feed["alpha"]) ** 2).mean()
        loss.backward()
        self.opt.step()
        return loss

    def clip_ims(self, ims):
        return np.uint8(np.clip((ims + 1) / 2 * 255, 0, 255))

    def save_multi_models(self, path_w, path_gan, **kw):
        self.saved.append(path_w)


def _opt(tmp_path, **kw):
    d = dict(output_dir=str(tmp_path), epochs=2, num_samples=12, max_iters=None, trainEmbed=False, updateGAN=False, layers=None,
             no_content_loss=True, no_gan_loss=True, log_every=1, model_save_freq=1000)
    d.update(kw)
    return types.SimpleNamespace(**d)


def _graph_util():
    return types.SimpleNamespace(graph_input=lambda g, n, seed=0: {"z": np.random.RandomState(seed).randn(n, 16)})


def test_multi_attr_targets_are_clamped_deltas():
    org = torch.tensor([[0.2, 0.9], [0.5, 0.1]])
    delta = torch.tensor([[0.5, 0.5], [-0.7, -0.7]])
    target, eps = train.multi_attr_targets(org, delta)
    assert torch.allclose(target, torch.tensor([[0.7, 1.0], [0.0, 0.0]]))
    assert torch.allclose(eps, target - org) and float(eps[0, 1]) == float(torch.tensor(1.0) - org[0, 1])


def test_single_attribute_loop_artifacts_and_targets(tmp_path):
    g = _StubGraph(1)
    consts = types.SimpleNamespace(BATCH_SIZE=4)
    out = train.train_loop(g, _opt(tmp_path), consts, _graph_util(), ["Smiling"])
    assert out == str(tmp_path) and len(g.calls) == 2 * 3                      # 2 epochs x (12 // 4) iterations
    assert g.saved == [f"{out}/model_w_0", f"{out}/model_w_1", f"{out}/model_w_2_final"]
    log = open(os.path.join(out, "log.txt")).read()
    assert log.count("T, epc, bst, lss, alpha:") == 6 and not os.path.exists(os.path.join(out, "loss_values.npy"))
    # train.py semantics: the sampled value itself is the regression target (one value for the whole batch)
    a = g.calls[0]["alpha"]
    assert a.shape == (4, 1) and float(a.min()) == float(a.max())


def test_multi_attr_loop_targets_losses_and_rank_sharding(tmp_path):
    g = _StubGraph(2)
    consts = types.SimpleNamespace(BATCH_SIZE=3)
    out = train.train_loop(g, _opt(tmp_path, epochs=3), consts, _graph_util(), ["night", "dark"], multi_attr=True)
    losses = np.load(os.path.join(out, "loss_values.npy"))
    assert losses.shape == (3 * 4,) and np.isfinite(losses).all()
    assert "alpha night:" in open(os.path.join(out, "log.txt")).read()
    for c in g.calls:                                                            # clamped per-sample targets
        assert c["alpha"].shape == (3, 2) and float(c["alpha"].min()) >= 0.0 and float(c["alpha"].max()) <= 1.0
    # rank r of a 2-rank job takes rows [i*6 + 3r, i*6 + 3r + 3) of the same seeded z: its first batch is the single
    # process's second batch, and both ranks draw identical targets
    g1 = _StubGraph(2)
    train.train_loop(g1, _opt(tmp_path / "r1", epochs=1), consts, _graph_util(), ["night", "dark"], rank=1, world=2, multi_attr=True)
    assert len(g1.calls) == 12 // 6
    assert torch.equal(g1.calls[0]["org"], g.calls[1]["org"])


def test_regressor_batchnorm_folding_preserves_values_and_input_gradients():
    """The frozen eval-mode ResNet-50 regressor with its BatchNorms folded into the convolutions (latent2im_b200/regressor.py)
    computes the same attribute predictions and the same gradient w.r.t. the image as the stock module."""
    import torch
    import torchvision
    from latent2im_b200.regressor import fold_batchnorm
    torch.manual_seed(0)
    m = torchvision.models.resnet50(weights=None)
    m.fc = torch.nn.Linear(2048, 40)
    m.eval()
    for mod in m.modules():          # non-trivial running statistics / affine parameters
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.1)
            mod.running_var.uniform_(0.5, 1.5)
            mod.weight.data.uniform_(0.5, 1.5)
            mod.bias.data.normal_(0, 0.1)
    f = fold_batchnorm(m)
    assert not any(isinstance(q, torch.nn.BatchNorm2d) for q in f.modules())
    x = torch.randn(2, 3, 64, 64)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = torch.sigmoid(m(xa)), torch.sigmoid(f(xb))
    assert (ya - yb).abs().max().item() <= 1e-4
    ya[:, 31].sum().backward()
    yb[:, 31].sum().backward()
    assert (xa.grad - xb.grad).abs().max().item() <= 1e-3 * xa.grad.abs().max().item()
    import pytest
    with pytest.raises(RuntimeError):
        fold_batchnorm(m.train())
