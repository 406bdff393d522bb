"""float64 models of the index algebra of the round-2 kernels, checked on CPU against direct formulations (no GPU, no oracle):
the 2x2-block GEMM of conv_tc_quad.cu with its structural zeros, and the rotating in-place FIR state of conv_tc_uprow.cu."""
import numpy as np
import torch
import torch.nn.functional as F


def _quad_matrix(w):
    """pack_quad_weight_kernel (conv_tc_quad.cu): Wq[n = (a*2+b)*32 + co][k = (r*4 + c)*32 + ci] = W[co][ci][r-a][c-b] or 0."""
    co_n, ci_n = w.shape[:2]
    wq = w.new_zeros(4 * co_n, 16 * ci_n)
    for a in range(2):
        for b in range(2):
            for r in range(4):
                for c in range(4):
                    kh, kw = r - a, c - b
                    if 0 <= kh < 3 and 0 <= kw < 3:
                        wq[(a * 2 + b) * co_n:(a * 2 + b + 1) * co_n, (r * 4 + c) * ci_n:(r * 4 + c + 1) * ci_n] = w[:, :, kh, kw]
    return wq


def test_quad_block_gemm_equals_conv_and_outer_patch_rows_are_half_empty():
    """One GEMM row = a 2x2 block of output pixels reading its 4x4 input patch (K = 16 * Cin).  Patch row 0 only reaches the upper
    pixel row (a = 0: N rows 0..2*Cout), patch row 3 only the lower one - the halves the kernel issues as N = 64 MMAs from
    half-size weight atoms must be exactly the non-zero ones."""
    torch.manual_seed(0)
    cin = cout = 8
    H = W = 6
    x = torch.randn(1, cin, H, W, dtype=torch.float64)
    w = torch.randn(cout, cin, 3, 3, dtype=torch.float64)
    ref = F.conv2d(x, w, padding=1)
    wq = _quad_matrix(w)
    xp = F.pad(x, (1, 1, 1, 1))
    out = torch.zeros_like(ref)
    for by in range(H // 2):
        for bx in range(W // 2):
            patch = xp[0, :, 2 * by:2 * by + 4, 2 * bx:2 * bx + 4]            # [ci][r][c]
            k = patch.permute(1, 2, 0).reshape(-1)                            # k = (r*4 + c)*Cin + ci
            n = wq @ k
            out[0, :, 2 * by:2 * by + 2, 2 * bx:2 * bx + 2] = n.reshape(2, 2, cout).permute(2, 0, 1)
    assert (out - ref).abs().max() < 1e-12
    half = 2 * cout
    k_row = 4 * cin                                                           # K extent of one patch row
    assert wq[half:, 0 * k_row:1 * k_row].abs().max() == 0                    # patch row 0 never feeds a = 1
    assert wq[:half, 3 * k_row:4 * k_row].abs().max() == 0                    # patch row 3 never feeds a = 0
    assert wq[:half, 0 * k_row:1 * k_row].abs().max() > 0 and wq[half:, 3 * k_row:4 * k_row].abs().max() > 0


def test_rotating_fir_state_with_folded_bias_equals_direct_fir():
    """conv_tc_uprow.cu epilogue: three partial output rows live in register sets that rotate with the row index (every update in
    place), the bias enters once per output row through the newest partial (pc' = f0 X + b/d):
    y[k-3] = D * (sum_i f[i] Hb[k-3+i] + b/d) + noise, finished when row k lands."""
    rng = np.random.RandomState(1)
    f = np.array([0.25, 0.75, 0.75, 0.25])
    nrows, width = 23, 5
    hb = rng.randn(nrows, width)
    d, bias = 1.7, -0.3
    bq = bias / d
    st = [np.zeros(width) for _ in range(3)]
    # rows before the run's first count as absent (the kernel writes output rows from k = 3 on)
    got = {}
    for k in range(nrows):
        ra, rb, rc = k % 3, (k + 1) % 3, (k + 2) % 3
        x = hb[k]
        if k >= 3:
            got[k - 3] = d * (f[3] * x + st[ra])
        st[rb] = f[2] * x + st[rb]
        st[rc] = f[1] * x + st[rc]
        st[ra] = f[0] * x + bq
    for j in range(nrows - 3):
        want = d * (f[0] * hb[j] + f[1] * hb[j + 1] + f[2] * hb[j + 2] + f[3] * hb[j + 3]) + bias
        assert np.abs(got[j] - want).max() < 1e-12, j
    # two-set variant (128 -> 64 / 256 -> 128 layers): pb' = f1 X + f0 X_prev + b/d re-reads the previous row
    st2 = [np.zeros(width) for _ in range(2)]
    got2 = {}
    for k in range(nrows):
        ra, rb = k % 2, (k + 1) % 2
        x, x1 = hb[k], (hb[k - 1] if k > 0 else np.zeros(width))
        if k >= 3:
            got2[k - 3] = d * (f[3] * x + st2[ra])
        st2[rb] = f[2] * x + st2[rb]
        st2[ra] = f[1] * x + (f[0] * x1 + bq)
    for j in range(nrows - 3):
        assert np.abs(got2[j] - got[j]).max() < 1e-12, j
