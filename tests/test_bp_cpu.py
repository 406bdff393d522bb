"""Host logic of BP.py on CPU: folder loading, the Gram perceptual term, the optimisation loop and the artefacts, driven
through a small differentiable stand-in generator (the real one is covered by tests/test_gpu_inversion.py)."""
import os

import numpy as np
import torch
from PIL import Image

import BP
from latent2im_b200.inversion import GramPerceptualLoss, invert, reconstruction_loss


class _StubGenerator(torch.nn.Module):
    requires_cuda = False
    n_latent = 4

    def __init__(self, size=16, dim=8):
        super().__init__()
        g = torch.Generator().manual_seed(0)
        self.basis = torch.nn.Parameter(torch.randn(dim, 3 * size * size, generator=g) / dim ** 0.5, requires_grad=False)
        self.size, self.dim = size, dim

    def mean_latent(self, n):
        return torch.zeros(1, self.dim)

    def forward(self, w, input_is_latent=True, noise=None):
        return torch.tanh(w.mean(1) @ self.basis).reshape(-1, 3, self.size, self.size), None


def _folder(tmp_path, n=3, size=20):
    d = tmp_path / "imgs" / "a"
    d.mkdir(parents=True)
    rs = np.random.RandomState(0)
    for i in range(n):
        Image.fromarray(rs.randint(0, 255, (size, size + 4, 3), dtype=np.uint8)).save(d / f"{i}.png")
    return str(tmp_path / "imgs")


def test_folder_loading_is_resize_crop_normalise(tmp_path):
    data = BP.load_image_folder(_folder(tmp_path), 16)
    x, label = data[0]
    assert len(data) == 3 and x.shape == (3, 16, 16) and label == 0 and -1.0 <= float(x.min()) and float(x.max()) <= 1.0


def test_gram_perceptual_loss_matches_formula():
    torch.manual_seed(0)
    loss = GramPerceptualLoss()
    a, b = torch.randn(2, 3, 32, 32), torch.randn(2, 3, 32, 32)
    got = loss(a, b)
    assert got.shape == (2,) and float(loss(a, a).abs().max()) == 0.0
    want = torch.zeros(2)
    x, y = a, b
    for i, m in enumerate(loss.features):                      # BP.py:173-185 written out
        x, y = m(x), m(y)
        if i in (3, 8, 15, 22):
            c, hw = x.shape[1], x.shape[2] * x.shape[3]
            gx = torch.einsum("bci,bdi->bcd", x.reshape(2, c, hw), x.reshape(2, c, hw)) / (c * hw)
            gy = torch.einsum("bci,bdi->bcd", y.reshape(2, c, hw), y.reshape(2, c, hw)) / (c * hw)
            want += ((gy - gx) ** 2).sum((1, 2)) * c * c
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-6)
    b.requires_grad_(True)
    loss(b, a).sum().backward()
    assert torch.isfinite(b.grad).all() and float(b.grad.abs().max()) > 0


def test_bp_run_writes_reference_artifacts_and_reduces_loss(tmp_path):
    args = BP.build_parser().parse_args(["--path", _folder(tmp_path), "--save_path", str(tmp_path / "out"), "--resolution", "32",
                                         "--batch_size", "2", "--n_loops", "40", "--lr", "0.05", "--optim", "Adam"])
    args.resolution = 16
    out = BP.run(args, generator=_StubGenerator(16), device=torch.device("cpu"))
    for name in ("org_0.png", "org_1.png", "0_final.png", "1_final.png", "latent/0_w.npy", "latent/1_w.npy", "loss_back.npy"):
        assert os.path.exists(os.path.join(out, name)), name
    assert Image.open(os.path.join(out, "org_0.png")).size == (16, 32)          # two images stacked vertically
    assert np.load(os.path.join(out, "latent", "0_w.npy")).shape == (2, 4, 8)
    losses = np.load(os.path.join(out, "loss_back.npy"))
    assert losses.shape == (80,) and losses[39] < losses[0] and losses[79] < losses[40]


def test_invert_objective_is_sum_of_squared_error_over_pixels():
    g = _StubGenerator(16)
    target = torch.rand(2, 3, 16, 16) * 2 - 1
    w, losses = invert(g, target, n_loops=1, lr=0.0, mean_latent=torch.zeros(1, 8))
    out, _ = g(w)
    assert torch.allclose(losses[0], reconstruction_loss(out, target).sum())
