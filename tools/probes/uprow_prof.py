"""Cycle accounting of conv_tc_uprow.cu (profiling build: -DL2I_UPROW_PROF, latent2im_b200/lib/libl2i_b200_uprowprof.so).
    L2I_LIB=latent2im_b200/lib/libl2i_b200_uprowprof.so python tools/probes/uprow_prof.py
Prints, per kernel variant (KC = Cin / 64) and warp role, the share of the kernel's clocks spent in each wait."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("L2I_ALLOW_RANDOM_INIT", "1")

import numpy as np
import torch


def main():
    from latent2im_b200 import _native as nt
    from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
    from latent2im_b200.synthetic import load_synthetic, synthetic_z
    dev = torch.device("cuda")
    b = 32
    gen = load_synthetic(Generator(1024, 512, 8), seed=0).to(dev).eval()
    gen.set_native(dtype=torch.bfloat16, max_batch=b)
    z = torch.tensor(synthetic_z(b, 0), dtype=torch.float32, device=dev)
    with torch.no_grad():
        lat = gen.style(z)[:, None, :].repeat(1, gen.n_latent, 1)
        for _ in range(3):
            gen(lat, input_is_latent=True)
    torch.cuda.synchronize()
    h = gen._handle(dev, b)
    n = 5 * 148 * 20 * 8
    buf = np.zeros(n, dtype=np.int64)
    fn = h.lib.l2i_debug_read_rgb_part
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    assert fn(h.handle, buf.ctypes.data, buf.nbytes) == 0
    buf = buf.reshape(5, 148, 20, 8)
    names_mma = ["issue", "wait tmem_empty", "wait a_full", "wait w_full", "-", "-", "-", "total"]
    names_epi = ["work", "wait tmem_full", "wait noise", "tcgen05.ld", "wait prev TMA store", "proxy fence", "-", "total"]
    for kc, layer in ((4, "convs.10 256->128"), (2, "convs.12 128->64"), (1, "convs.14 64->32")):
        blk = buf[kc].astype(np.float64)
        if blk[:, 1, 7].sum() == 0:
            continue
        print(f"--- KC={kc} ({layer}) ---")
        tot = blk[:, 1, 7].mean()
        print(f"  MMA issuer: total {tot:.0f} clk/CTA; " + ", ".join(f"{nm} {100 * blk[:, 1, i].mean() / tot:.1f}%" for i, nm in enumerate(names_mma[:4])))
        for w in range(4, 12):
            tot = blk[:, w, 7].mean()
            print(f"  epilogue warp {w} (quadrant {w & 3}, channel half {(w - 4) >> 2}): total {tot:.0f}; " +
                  ", ".join(f"{nm} {100 * blk[:, w, i].mean() / tot:.1f}%" for i, nm in enumerate(names_epi[:6])))
        e = blk[:, 4:, :]
        print("  slowest-vs-fastest epilogue warp 'work' clocks per CTA (mean over CTAs): "
              f"{e[:, :, 0].max(1).mean():.0f} vs {e[:, :, 0].min(1).mean():.0f}")


if __name__ == "__main__":
    main()
