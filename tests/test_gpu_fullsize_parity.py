"""GPU: parity at the BASELINE.json sizes (cfg1 256 px, cfg2 1024 px) against (i) fixtures the UNMODIFIED reference produced
on a B200 (tests/golden/ref_gpu_fullsize.npz, make_golden_ref_gpu_fullsize.py) and (ii) the float64 CPU oracle, forward and
latent gradient, fp32 CUDA-core path and bf16 tcgen05 path, on the UNSCALED SURVEY 8d recipe (unit-variance ToRGB weights).

Every kernel that runs at res >= 256 (A-resident, halo-resident, vertical-pair, 2x2-block, row-marching fused up-conv, the
tcgen05 data-gradient convs with strided TMA boxes) is therefore compared with the reference here, not with the repo's own
fp32 path.  Gates: fp32 max-abs <= 1e-3 (north star); bf16 PSNR >= 45 dB at 256 px; at 1024 px the synthetic image of this
seed spans 22 (a trained generator's spans 2), so the peak-to-peak-2 PSNR is reported and gated at the measured level while
the amplitude-independent form (peak = the reference image's own range) is gated at 55 dB."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden_ref_gpu_fullsize import CASES, CROP, CROPS_1024, fullsize_inputs  # noqa: E402

from latent2im_b200.synthetic import load_synthetic, synthetic_z  # noqa: E402

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_gpu_fullsize.npz")


def _case(tag):
    (size, batch, seed), = [(s, b, sd) for t, s, b, sd in CASES if t == tag]
    return size, batch, seed


def _run(tag, dtype, grad=False):
    """Product path on the fixture's inputs; returns (image, latent gradient or None, mapping output)."""
    from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
    size, batch, seed = _case(tag)
    gen = load_synthetic(Generator(size, 512, 8), seed=seed).cuda()          # rgb_gain 1.0: the unscaled recipe
    gen.set_native(dtype=dtype, max_batch=batch)
    z = torch.tensor(synthetic_z(batch, 10 + seed), dtype=torch.float32).cuda()
    with torch.no_grad():
        w = gen.style(z)
    lat, noise, probe = fullsize_inputs(size, batch, seed, gen.n_latent, gen.num_layers, w)
    glat = None
    if grad:
        lat.requires_grad_(True)
        img, _ = gen(lat, input_is_latent=True, noise=noise)
        (img * probe).sum().backward()
        glat = lat.grad.detach()
        img = img.detach()
    else:
        with torch.no_grad():
            img, _ = gen(lat, input_is_latent=True, noise=noise)
    return img, glat, w


def _psnr_pair(err_sq_mean, span):
    return 10 * np.log10(4.0 / err_sq_mean), 10 * np.log10(span ** 2 / err_sq_mean)


def _image_errors(tag, img, z):
    """max-abs and mean-square error of `img` against the stored parts of the reference image."""
    if f"{tag}_image" in z:
        d = img.cpu().double() - torch.from_numpy(z[f"{tag}_image"]).double()
        return d.abs().max().item(), (d ** 2).mean().item()
    crops = torch.from_numpy(z[f"{tag}_crops"]).double()
    d = torch.stack([img[:, :, y:y + CROP, x:x + CROP].cpu().double() for (y, x) in CROPS_1024]) - crops
    pooled = torch.nn.functional.avg_pool2d(img.double(), 4).cpu() - torch.from_numpy(z[f"{tag}_pooled4"]).double()
    return max(d.abs().max().item(), pooled.abs().max().item()), (d ** 2).mean().item()


@pytest.mark.parametrize("tag", ["s256", "s1024"])
def test_fp32_path_matches_reference_fixture(tag):
    z = np.load(GOLD)
    img, glat, w = _run(tag, torch.float32, grad=True)
    assert (w.cpu() - torch.from_numpy(z[f"{tag}_w"])).abs().max().item() <= 1e-4
    mom = z[f"{tag}_moments"]
    span = float(mom[3].max() - mom[2].min())
    max_abs, _ = _image_errors(tag, img, z)
    assert max_abs <= 1e-3, f"{tag}: fp32 max-abs {max_abs:.3e} (image spans {span:.1f})"
    gref = torch.from_numpy(z[f"{tag}_grad_latent"])
    # leaky-relu kinks make single gradient components differ by ~1e-3 of the maximum between two fp32 implementations
    # (an activation within rounding of 0 takes the other slope): gate the relative L2 error, bound the max norm loosely
    gerr = (glat.cpu() - gref).abs().max().item() / gref.abs().max().item()
    gl2 = ((glat.cpu().double() - gref.double()).norm() / gref.double().norm()).item()
    assert gl2 <= 4e-3 and gerr <= 2e-2, f"{tag}: fp32 latent gradient relative L2 error {gl2:.3e}, max {gerr:.3e}"


@pytest.mark.parametrize("tag,gate_p2p2", [("s256", 45.0), ("s1024", 40.0)])
def test_bf16_path_matches_reference_fixture(tag, gate_p2p2):
    z = np.load(GOLD)
    img, glat, _ = _run(tag, torch.bfloat16, grad=True)
    assert torch.isfinite(img).all()
    mom = z[f"{tag}_moments"]
    span = float(mom[3].max() - mom[2].min())
    _, mse = _image_errors(tag, img, z)
    p2, own = _psnr_pair(mse, span)
    print(f"{tag}: bf16 forward vs reference fixture: PSNR(p2p 2) {p2:.2f} dB, PSNR(own range {span:.1f}) {own:.2f} dB")
    assert p2 >= gate_p2p2 and own >= 55.0, (tag, p2, own)
    # bf16 data-gradient kernels (training-mode forward + backward) against the reference's fp32 autograd gradient
    gref = torch.from_numpy(z[f"{tag}_grad_latent"]).double().flatten()
    g = glat.cpu().double().flatten()
    cos = torch.nn.functional.cosine_similarity(g, gref, dim=0).item()
    rel = ((g - gref).norm() / gref.norm()).item()
    print(f"{tag}: bf16 latent gradient vs reference: cosine {cos:.5f}, relative L2 error {rel:.4f}")
    assert cos >= 0.995 and rel <= 0.10, (tag, cos, rel)


def test_inference_forward_matches_training_forward_bf16():
    """The no-grad forward (fused up-conv kernels, uint8-capable last layer) and the training-mode forward (two-kernel
    up-conv keeping t) are different kernel sets: both must agree with the reference fixture at 1024 px."""
    z = np.load(GOLD)
    img_inf, _, _ = _run("s1024", torch.bfloat16, grad=False)
    img_trn, _, _ = _run("s1024", torch.bfloat16, grad=True)
    mom = z["s1024_moments"]
    span = float(mom[3].max() - mom[2].min())
    for img in (img_inf, img_trn):
        _, mse = _image_errors("s1024", img, z)
        assert _psnr_pair(mse, span)[1] >= 55.0


@pytest.mark.parametrize("dtype,gate", [(torch.float32, None), (torch.bfloat16, 55.0)])
def test_1024px_forward_matches_float64_oracle(dtype, gate):
    """cfg2 size, batch 1, every pixel: the product path against the float64 CPU restatement (seconds on the box's cores)."""
    from oracle import GeneratorSpec, generator_forward_ref
    from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    size, batch, seed = _case("s1024")
    gen = load_synthetic(Generator(size, 512, 8), seed=seed)
    sd = {k: v.double() for k, v in gen.state_dict().items()}
    img, _, w = _run("s1024", dtype)
    spec = GeneratorSpec(size=size)
    lat, noise, _ = fullsize_inputs(size, batch, seed, spec.n_latent, spec.num_layers, w.cpu())
    ref = generator_forward_ref(sd, lat.double(), [n.double() for n in noise], spec)
    d = img.cpu().double() - ref
    span = (ref.max() - ref.min()).item()
    if gate is None:
        assert d.abs().max().item() <= 1e-3, d.abs().max().item()
    else:
        p2, own = _psnr_pair((d ** 2).mean().item(), span)
        print(f"1024 px bf16 vs float64 oracle: PSNR(p2p 2) {p2:.2f} dB, PSNR(own range {span:.1f}) {own:.2f} dB")
        assert own >= gate
