"""GPU: walk modules (drop-in classes over the native walk kernels) vs the CPU oracle, incl. gradients."""
import numpy as np
import pytest
import torch

from oracle.walks import walk_linear_ref, walk_mlp_ref, walk_nonlinear_ref

pytestmark = pytest.mark.gpu


def _mods():
    from latent2im_b200.graphs.stylegan_v2_real import transform_base as tb
    return tb


def _seq_params(seq):
    return [(m.weight.detach().cpu().double(), m.bias.detach().cpu().double()) for m in seq if isinstance(m, torch.nn.Linear)]


@pytest.mark.parametrize("shared", [True, False])
@pytest.mark.parametrize("layers", [None, [0, 3]])
def test_walk_linear_forward_and_grad(shared, layers):
    tb = _mods()
    np.random.seed(0)
    walk = tb.WalkLinearMultiW(64, 2, 1, ["a", "b"]).cuda()
    B = 3
    base = torch.randn(B, 64, device="cuda")
    ws = [base] * 6 if shared else [torch.randn(B, 64, device="cuda") for _ in range(6)]
    alpha = torch.randn(B, 2, device="cuda")
    out = walk(ws, alpha, layers=layers)
    ref = walk_linear_ref([w.cpu().double() for w in ws], alpha.cpu().double(), walk.w.detach().cpu().double(), layers)
    for a, b in zip(out, ref):
        assert torch.allclose(a.detach().cpu().double(), b, atol=1e-5)
    probe = torch.randn(B, 6, 64, device="cuda")
    (torch.stack(out, 1) * probe).sum().backward()
    wr = walk.w.detach().cpu().double().requires_grad_(True)
    ref2 = walk_linear_ref([w.cpu().double() for w in ws], alpha.cpu().double(), wr, layers)
    (torch.stack(ref2, 1) * probe.cpu().double()).sum().backward()
    assert torch.allclose(walk.w.grad.cpu().double(), wr.grad, atol=1e-4)


def test_walk_mlp_forward_and_grad():
    tb = _mods()
    torch.manual_seed(0)
    walk = tb.WalkMlpMultiW(32, 1, 1, ["a"]).cuda()
    base = torch.randn(4, 32, device="cuda")
    ws = [base] * 4
    alpha = torch.randn(4, 1, device="cuda")
    params = _seq_params(walk.linear)
    with torch.no_grad():
        out = walk(ws, alpha)
    ref = walk_mlp_ref([w.cpu().double() for w in ws], alpha.cpu().double(), params)
    for a, b in zip(out, ref):
        assert torch.allclose(a.cpu().double(), b, atol=1e-4)
    with torch.no_grad():
        out = walk(ws, alpha, layers=[1, 2])
    ref = walk_mlp_ref([w.cpu().double() for w in ws], alpha.cpu().double(), params, layers=[1, 2])
    for a, b in zip(out, ref):
        assert torch.allclose(a.cpu().double(), b, atol=1e-4)
    # gradients w.r.t. the MLP parameters
    out = walk(ws, alpha)
    probe = torch.randn(4, 4, 32, device="cuda")
    (torch.stack(out, 1) * probe).sum().backward()
    pr = [(w.clone().requires_grad_(True), b.clone().requires_grad_(True)) for w, b in params]
    ref = walk_mlp_ref([w.cpu().double() for w in ws], alpha.cpu().double(), pr)
    (torch.stack(ref, 1) * probe.cpu().double()).sum().backward()
    lin = [m for m in walk.linear if isinstance(m, torch.nn.Linear)]
    for m, (w, b) in zip(lin, pr):
        assert torch.allclose(m.weight.grad.cpu().double(), w.grad, atol=1e-3, rtol=1e-3)
        assert torch.allclose(m.bias.grad.cpu().double(), b.grad, atol=1e-3, rtol=1e-3)


def test_walk_nonlinear_forward():
    tb = _mods()
    torch.manual_seed(1)
    walk = tb.WalkNonLinearW(32, 1, 1, ["a"]).cuda()
    ws = [torch.randn(3, 32, device="cuda") for _ in range(4)]
    alpha = torch.randn(3, 1, device="cuda")
    emb = (walk.embed.weight.detach().cpu().double(), walk.embed.bias.detach().cpu().double())
    params = _seq_params(walk.linear)
    with torch.no_grad():
        out = walk(ws, alpha=alpha)                       # the keyword call get_w_new_tensor makes
        out_l = walk(ws, None, alpha, None, layers=[0, 2])  # the reference's positional signature
    ref = walk_nonlinear_ref([w.cpu().double() for w in ws], alpha.cpu().double(), emb, params)
    ref_l = walk_nonlinear_ref([w.cpu().double() for w in ws], alpha.cpu().double(), emb, params, layers=[0, 2])
    for a, b in zip(out, ref):
        assert torch.allclose(a.cpu().double(), b, atol=1e-4)
    for a, b in zip(out_l, ref_l):
        assert torch.allclose(a.cpu().double(), b, atol=1e-4)


@pytest.mark.parametrize("shared", [True, False])
@pytest.mark.parametrize("layers", [None, [0, 2]])
def test_walk_nonlinear_gradients(shared, layers):
    """Training path of WalkNonLinearW (embed -> MLP -> d / ||d|| -> add) runs on l2i_linear_fwd/bwd and
    l2i_walk_combine/_bwd only; parameter AND input-latent gradients vs float64 autograd through the oracle
    (transform_base.py:219-243).  layers=None normalises d, a layer list does not (reference behaviour)."""
    tb = _mods()
    torch.manual_seed(3)
    walk = tb.WalkNonLinearW(32, 1, 1, ["a"]).cuda()
    B, n = 5, 4
    base = torch.randn(B, 32, device="cuda")
    ws = [base.clone().requires_grad_(True)] * n if shared else [torch.randn(B, 32, device="cuda").requires_grad_(True) for _ in range(n)]
    alpha = torch.randn(B, 1, device="cuda")
    probe = torch.randn(B, n, 32, device="cuda")
    out = walk(ws, alpha=alpha, layers=layers)
    (torch.stack(out, 1) * probe).sum().backward()

    emb = tuple(t.detach().cpu().double().requires_grad_(True) for t in (walk.embed.weight, walk.embed.bias))
    pr = [(w.clone().requires_grad_(True), b.clone().requires_grad_(True)) for w, b in _seq_params(walk.linear)]
    if shared:
        b64 = ws[0].detach().cpu().double().requires_grad_(True)
        ws64 = [b64] * n
    else:
        ws64 = [w.detach().cpu().double().requires_grad_(True) for w in ws]
    ref = walk_nonlinear_ref(ws64, alpha.cpu().double(), emb, pr, layers=layers)
    (torch.stack(ref, 1) * probe.cpu().double()).sum().backward()
    for a, b in zip(out, ref):
        assert torch.allclose(a.detach().cpu().double(), b.detach(), atol=1e-4)
    lin = [m for m in walk.linear if isinstance(m, torch.nn.Linear)]
    for m, (w, b) in zip([walk.embed] + lin, [emb] + pr):
        assert torch.allclose(m.weight.grad.cpu().double(), w.grad, atol=2e-3, rtol=2e-3), m
        assert torch.allclose(m.bias.grad.cpu().double(), b.grad, atol=2e-3, rtol=2e-3), m
    got_in = [ws[0].grad] if shared else [w.grad for w in ws]
    ref_in = [ws64[0].grad] if shared else [w.grad for w in ws64]
    for a, b in zip(got_in, ref_in):
        assert torch.allclose(a.cpu().double(), b, atol=2e-3, rtol=2e-3)


def test_walk_mlp_gradients_distinct_inputs_and_layers():
    """WalkMlpMultiW with a genuinely per-layer W+ list (one MLP evaluation per layer, stacked d) restricted to two layers."""
    tb = _mods()
    torch.manual_seed(4)
    walk = tb.WalkMlpMultiW(32, 1, 1, ["a"]).cuda()
    B, n = 3, 4
    ws = [torch.randn(B, 32, device="cuda").requires_grad_(True) for _ in range(n)]
    alpha = torch.randn(B, 1, device="cuda")
    probe = torch.randn(B, n, 32, device="cuda")
    out = walk(ws, alpha, layers=[1, 3])
    (torch.stack(out, 1) * probe).sum().backward()
    pr = [(w.clone().requires_grad_(True), b.clone().requires_grad_(True)) for w, b in _seq_params(walk.linear)]
    ws64 = [w.detach().cpu().double().requires_grad_(True) for w in ws]
    ref = walk_mlp_ref(ws64, alpha.cpu().double(), pr, layers=[1, 3])
    (torch.stack(ref, 1) * probe.cpu().double()).sum().backward()
    lin = [m for m in walk.linear if isinstance(m, torch.nn.Linear)]
    for m, (w, b) in zip(lin, pr):
        assert torch.allclose(m.weight.grad.cpu().double(), w.grad, atol=1e-3, rtol=1e-3)
        assert torch.allclose(m.bias.grad.cpu().double(), b.grad, atol=1e-3, rtol=1e-3)
    for a, b in zip(ws, ws64):
        assert torch.allclose(a.grad.cpu().double(), b.grad, atol=1e-3, rtol=1e-3)


def test_walk_linear_rejects_mismatched_parameter_shape():
    """A walk trained for 14 latent layers applied to an 18-layer W+ list (or alpha with the wrong attribute count) must
    raise, not read past the parameter buffer."""
    tb = _mods()
    np.random.seed(0)
    walk = tb.WalkLinearMultiW(64, 2, 1, ["a", "b"]).cuda()       # w: [2, 6, 64]
    base = torch.randn(3, 64, device="cuda")
    with pytest.raises(RuntimeError, match="does not match"):
        walk([base] * 8, torch.randn(3, 2, device="cuda"))
    with pytest.raises(RuntimeError, match="does not match"):
        walk([base] * 6, torch.randn(3, 3, device="cuda"))


def test_walk_pickle_roundtrip_under_reference_module_path(tmp_path):
    import latent2im_b200
    latent2im_b200.install_dropin()
    from graphs.stylegan_v2_real.transform_base import WalkLinearMultiW  # reference import path
    walk = WalkLinearMultiW(16, 1, 1, ["a"])
    p = tmp_path / "model_w_0_walk_module.ckpt"
    torch.save(walk, p)
    back = torch.load(p, weights_only=False)
    assert torch.equal(back.w, walk.w)


def test_walk_trainer_step_matches_oracle_autograd():
    """One train.py-style step (G fwd no-grad, R, eps, walk, G fwd+bwd, BCE, Adam) through WalkTrainer:
    the walk-parameter gradient equals autograd through the float64 oracle with the same (tiny) regressor."""
    from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
    from latent2im_b200.graphs.stylegan_v2_real.transform_base import WalkLinearMultiW
    from latent2im_b200.synthetic import load_synthetic, synthetic_noise, synthetic_walk_w, synthetic_z
    from latent2im_b200.train_step import WalkTrainer, bce_clamped
    from oracle import GeneratorSpec, generator_forward_ref, mapping_ref
    from oracle.walks import walk_linear_ref

    size, dim, n_mlp, batch = 16, 32, 2, 3
    spec = GeneratorSpec(size=size, style_dim=dim, n_mlp=n_mlp)
    gen = load_synthetic(Generator(size, dim, n_mlp), seed=0)
    sd = {k: v.double() for k, v in gen.state_dict().items()}
    gen = gen.cuda().eval()
    gen.set_native(dtype=torch.float32, max_batch=batch)
    torch.manual_seed(1)
    reg = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.Tanh(), torch.nn.AdaptiveAvgPool2d(1),
                              torch.nn.Flatten(), torch.nn.Linear(4, 5), torch.nn.Sigmoid())
    reg64 = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.Tanh(), torch.nn.AdaptiveAvgPool2d(1),
                                torch.nn.Flatten(), torch.nn.Linear(4, 5), torch.nn.Sigmoid()).double()
    reg64.load_state_dict({k: v.double() for k, v in reg.state_dict().items()})
    reg = reg.cuda()
    walk = WalkLinearMultiW(dim, spec.log_size - 2, 1, ["Smiling"]).cuda()
    w0 = synthetic_walk_w(1, spec.n_latent, dim, seed=0)
    with torch.no_grad():
        walk.w.copy_(w0.cuda())
    z = torch.tensor(synthetic_z(batch, 0, dim), dtype=torch.float32)
    target = torch.full((batch, 1), 0.8)
    noise = synthetic_noise(spec.num_layers, batch)

    # product path, with explicit noise so that both G forwards are reproducible: patch the generator call
    trainer = WalkTrainer(gen, walk, reg, [2], lr=1e-3)
    orig_forward = gen.forward
    gen.forward = lambda styles, **kw: orig_forward(styles, noise=[n.cuda() for n in noise], **kw)
    loss = trainer.step(z.cuda(), target.cuda())
    g_native = None
    # the optimizer already stepped: recover the gradient from a fresh backward at the ORIGINAL parameters
    with torch.no_grad():
        walk.w.copy_(w0.cuda())
    walk.w.grad = None
    w = gen.style(z.cuda())
    with torch.no_grad():
        img0, _ = gen(w[:, None, :].expand(-1, spec.n_latent, -1), input_is_latent=True)
        eps = target.cuda() - reg(img0)[:, [2]]
    lat = torch.stack(walk([w] * spec.n_latent, eps), 1)
    img, _ = gen(lat, input_is_latent=True)
    loss2 = bce_clamped(reg(img)[:, [2]], target.cuda())
    loss2.backward()
    g_native = walk.w.grad.detach().cpu().double()
    assert abs(loss.item() - loss2.item()) <= 1e-5 * max(1.0, abs(loss2.item()))

    # oracle (float64, CPU autograd)
    wr = mapping_ref(sd, z.double(), spec)
    with torch.no_grad():
        img0r = generator_forward_ref(sd, wr[:, None, :].repeat(1, spec.n_latent, 1), noise, spec)
        epsr = target.double() - reg64(img0r)[:, [2]]
    wp = w0.double().clone().requires_grad_(True)
    latr = torch.stack(walk_linear_ref([wr] * spec.n_latent, epsr, wp), 1)
    imgr = generator_forward_ref(sd, latr, noise, spec)
    lossr = bce_clamped(reg64(imgr)[:, [2]], target.double())
    (g_ref,) = torch.autograd.grad(lossr, wp)
    assert abs(loss2.item() - lossr.item()) <= 1e-3 * max(1.0, abs(lossr.item()))
    scale = g_ref.abs().max().item()
    assert (g_native - g_ref).abs().max().item() <= 5e-3 * scale, ((g_native - g_ref).abs().max().item(), scale)
