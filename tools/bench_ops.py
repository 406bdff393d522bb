"""Achieved HBM bandwidth of the standalone drop-in ops (north-star item 3) at BASELINE cfg2 shapes, next to the measured copy
peak: l2i_upfirdn2d in the three modes the reference uses (Blur after the up-conv, Upsample of the skip image, its
transpose / Downsample), l2i_fused_bias_act forward and l2i_fused_leaky_relu_bwd.

    gpurun -- python tools/bench_ops.py [--json gpurun_out/ops.json]

Algorithmic bytes = read the input once + write the output once; timing = CUDA events over 20 launches after 5
warm-ups on tensors far larger than the 126 MB L2.  Measurement infrastructure only.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch


def timed(fn, steps=20, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    from latent2im_b200.graphs.stylegan_v2_real.op import fused_leaky_relu, upfirdn2d
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    peak = peaks["hbm_gbs"]
    dev = torch.device("cuda")
    k = torch.tensor([1., 3., 3., 1.], device=dev)
    k2 = k[None, :] * k[:, None]
    k2 = k2 / k2.sum()
    rows = []

    def rec(name, ms, nbytes, shape):
        gbs = nbytes / (ms * 1e-3) / 1e9
        rows.append({"op": name, "shape": shape, "ms": round(ms, 4), "algorithmic_mb": round(nbytes / 1e6, 1), "gbs": round(gbs, 1),
                     "frac_of_measured_hbm_peak": round(gbs / peak, 3)})
        print(rows[-1], flush=True)

    for dt, es in ((torch.float32, 4), (torch.bfloat16, 2)):
        tag = "fp32" if es == 4 else "bf16"
        # Blur after the stride-2 transposed conv (networks.py:72-88): [B, C, 2H+1, 2H+1] -> [B, C, 2H, 2H], pad (1, 1)
        for (b, c, h) in ((8, 32, 1025), (8, 64, 513)):
            x = torch.randn(b, c, h, h, device=dev, dtype=dt)
            ms = timed(lambda: upfirdn2d(x, k2 * 4, pad=(1, 1)))
            rec(f"upfirdn2d blur (up 1, down 1, pad 1,1) {tag}", ms, b * c * (h * h + (h - 1) ** 2) * es, [b, c, h, h])
            del x
        # Upsample of the skip image (networks.py:30-48): [B, 3, H, H] -> [B, 3, 2H, 2H], up 2, pad (2, 1)
        x = torch.randn(64, 3, 512, 512, device=dev, dtype=dt)
        ms = timed(lambda: upfirdn2d(x, k2 * 4, up=2, pad=(2, 1)))
        rec(f"upfirdn2d upsample (up 2, pad 2,1) {tag}", ms, 64 * 3 * (512 * 512 + 1024 * 1024) * es, [64, 3, 512, 512])
        del x
        # its transpose / Downsample: [B, 3, 2H, 2H] -> [B, 3, H, H], down 2, pad (1, 1)
        x = torch.randn(64, 3, 1024, 1024, device=dev, dtype=dt)
        ms = timed(lambda: upfirdn2d(x, k2, down=2, pad=(1, 1)))
        rec(f"upfirdn2d downsample (down 2, pad 1,1) {tag}", ms, 64 * 3 * (512 * 512 + 1024 * 1024) * es, [64, 3, 1024, 1024])
        del x
        # fused bias + leaky relu (op/fused_act.py:51-86) at the largest cfg2 activation slice that fits comfortably
        x = torch.randn(8, 32, 1024, 1024, device=dev, dtype=dt)
        bias = torch.randn(32, device=dev, dtype=dt)
        ms = timed(lambda: fused_leaky_relu(x, bias))
        rec(f"fused_bias_act forward {tag}", ms, 2 * x.numel() * es, list(x.shape))
        xg = x.clone().requires_grad_(True)
        bg = bias.clone().requires_grad_(True)
        y = fused_leaky_relu(xg, bg)
        gy = torch.randn_like(y)

        def bwd():
            torch.autograd.grad(y, (xg, bg), gy, retain_graph=True)

        ms = timed(bwd)
        rec(f"fused_leaky_relu backward (grad_in + grad_bias) {tag}", ms, 3 * x.numel() * es, list(x.shape))
        del x, xg, y, gy
        torch.cuda.empty_cache()
    out = {"peak_hbm_gbs": peak, "peak_source": "MEASURED_PEAKS.json" if "gpu_name" in peaks else "fallback", "rows": rows}
    if a.json:
        os.makedirs(os.path.dirname(os.path.abspath(a.json)), exist_ok=True)
        json.dump(out, open(a.json, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
