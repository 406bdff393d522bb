"""Float64 CPU model of the *fused kernel algorithm* (test infrastructure).

Follows latent2im_b200/csrc/{generator,conv_simt,pointwise}.cu step by step - pre-scaled
activations, shared weights with demodulation as an output scale, stride-2 transposed conv as
four output phases written into a padded (2H+2)^2 buffer, blur + noise + bias + lrelu + next-style
scale, ToRGB partial sums in the conv epilogue, fused skip up-sampling - so the algebra can be
checked against the oracle on CPU before any GPU time is spent.
"""
import math

import torch


def _taps_plain():
    return [(kh - 1, kw - 1, kh, kw) for kh in range(3) for kw in range(3)]


def _taps_up(py, px):
    return [(-(kh // 2), -(kw // 2), kh, kw) for kh in range(py, 3, 2) for kw in range(px, 3, 2)]


def _tap_conv(x, w, taps, oh, ow):
    """x: [B,H,W,Cin] (NHWC); w: [Cout,Cin,3,3]; out[b,oy,ox,co] = sum_t sum_ci w[co,ci,kh,kw] * x[b,oy+dy,ox+dx,ci]."""
    b, h, wd, cin = x.shape
    out = x.new_zeros(b, oh, ow, w.shape[0])
    for dy, dx, kh, kw in taps:
        shifted = x.new_zeros(b, oh, ow, cin)
        y0, y1 = max(0, -dy), min(oh, h - dy)
        x0, x1 = max(0, -dx), min(ow, wd - dx)
        if y1 > y0 and x1 > x0:
            shifted[:, y0:y1, x0:x1] = x[:, y0 + dy:y1 + dy, x0 + dx:x1 + dx]
        out = out + torch.einsum("bhwc,oc->bhwo", shifted, w[:, :, kh, kw])
    return out


def composite_weights(wt, f):
    """Mirrors pack_composite_weight_kernel (pointwise.cu): wt [Cout,Cin,3,3] (already scaled), f = flipped blur taps * 2 / sum.
    Returns Wc [4 phases][3 dy][3 dx][Cout][Cin]: transposed conv (stride 2) + blur as a 3x3 conv per output phase."""
    cout, cin = wt.shape[:2]
    wc = wt.new_zeros(4, 3, 3, cout, cin)
    for ph in range(4):
        py, px = ph >> 1, ph & 1
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                for i in range(4):
                    kh = py + i - 1 - 2 * dy
                    if kh < 0 or kh > 2:
                        continue
                    for j in range(4):
                        kw = px + j - 1 - 2 * dx
                        if kw < 0 or kw > 2:
                            continue
                        wc[ph, dy + 1, dx + 1] += f[i] * f[j] * wt[:, :, kh, kw]
    return wc


def _composite_upconv(x, wc, H):
    """x: [B,H,H,Cin] -> blur(convT(x)) [B,2H,2H,Cout] via the per-phase 3x3 convs at input resolution."""
    b = x.shape[0]
    cout = wc.shape[3]
    out = x.new_zeros(b, 2 * H, 2 * H, cout)
    for ph in range(4):
        py, px = ph >> 1, ph & 1
        acc = x.new_zeros(b, H, H, cout)
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                shifted = x.new_zeros(x.shape)
                y0, y1 = max(0, -dy), min(H, H - dy)
                x0, x1 = max(0, -dx), min(H, H - dx)
                shifted[:, y0:y1, x0:x1] = x[:, y0 + dy:y1 + dy, x0 + dx:x1 + dx]
                acc = acc + torch.einsum("bhwc,oc->bhwo", shifted, wc[ph, dy + 1, dx + 1])
        out[:, py::2, px::2] = acc
    return out


def rowfold_weights(wt, f):
    """Mirrors pack_uprow_weight_kernel (conv_tc_uprow.cu): only the HORIZONTAL blur is folded into the weights.
    H[u][2n+b] = sum_j f[j] t[u][2n+b-1+j],  t[u][v] = sum_kw W[kh][kw] x[.][(v-kw)/2]  =>  per (kh, dx = -1..1, b):
    Wr[kh][dx][b] = sum_{j : kw = b + j - 1 - 2 dx in [0,2]} f[j] W[kh][kw].   Returns [3 kh][3 dx][2 b][Cout][Cin]."""
    cout, cin = wt.shape[:2]
    wr = wt.new_zeros(3, 3, 2, cout, cin)
    for kh in range(3):
        for dx in (-1, 0, 1):
            for b in range(2):
                for j in range(4):
                    kw = b + j - 1 - 2 * dx
                    if 0 <= kw <= 2:
                        wr[kh, dx + 1, b] += f[j] * wt[:, :, kh, kw]
    return wr


def _rowfold_upconv(x, wr, f, H):
    """Row-marching fused up-conv (conv_tc_uprow.cu): horizontally blurred transposed-conv rows H[u], u = -1 .. 2H+1
    (u = 2m + py gathers kh = py (mod 2) from input row m - kh/2, zero outside), then the vertical 4-tap FIR
    out[oy] = sum_i f[i] H[oy - 1 + i].   x: [B,H,H,Cin] -> [B,2H,2H,Cout]."""
    b_, _, W, cin = x.shape
    cout = wr.shape[3]
    xp = x.new_zeros(b_, H + 2, W + 2, cin)          # zero border = TMA out-of-bounds fill
    xp[:, 1:H + 1, 1:W + 1] = x
    rows = {}
    for u in range(-1, 2 * H + 2):
        acc = x.new_zeros(b_, W, 2, cout)
        for kh in range(3):
            if (u - kh) % 2:
                continue
            m = (u - kh) // 2                         # input row; -1 and H read the zero border
            if m < -1 or m > H:
                continue
            for dx in (-1, 0, 1):
                acc = acc + torch.einsum("bwc,poc->bwpo", xp[:, m + 1, 1 + dx:1 + dx + W], wr[kh, dx + 1])
        rows[u] = acc.reshape(b_, 2 * W, cout)
    out = x.new_zeros(b_, 2 * H, 2 * W, cout)
    for oy in range(2 * H):
        for i in range(4):
            out[:, oy] += f[i] * rows[oy - 1 + i]
    return out


def _upsample2x(skip, f):
    """skip: [B,3,h,w] -> [B,3,2h,2w]; mirrors upsample2x_at (conv_common.cuh)."""
    b, c, h, w = skip.shape
    out = skip.new_zeros(b, c, 2 * h, 2 * w)
    for Y in range(2 * h):
        for i in range(4):
            u = Y + i - 2
            if u < 0 or u % 2 or u // 2 >= h:
                continue
            for X in range(2 * w):
                for j in range(4):
                    v = X + j - 2
                    if v < 0 or v % 2 or v // 2 >= w:
                        continue
                    out[:, :, Y, X] += f[i] * f[j] * skip[:, :, u // 2, v // 2]
    return out


def fused_forward_model(sd, latent, noise, spec, dtype=torch.float64, composite=False):
    """Returns (image, {name: unscaled activation NCHW}, {k: skip})."""
    sd = {k: v.to(dtype) for k, v in sd.items()}
    latent = latent.to(dtype)
    B, D = latent.shape[0], spec.style_dim
    taps = torch.tensor(spec.blur_taps, dtype=dtype)
    f = torch.flip(taps, [0]) / taps.sum() * 2

    # layer table
    convs = [("conv1", False, 0, 0)]
    rgbs = [("to_rgb1", 1)]
    for k in range(spec.log_size - 2):
        convs.append((f"convs.{2 * k}", True, 2 * k + 1, 2 * k + 1))
        convs.append((f"convs.{2 * k + 1}", False, 2 * k + 2, 2 * k + 2))
        rgbs.append((f"to_rgbs.{k}", 2 * k + 3))

    def style(prefix, li):
        w = sd[prefix + ".conv.modulation.weight"] * (1 / math.sqrt(D))
        return latent[:, li] @ w.t() + sd[prefix + ".conv.modulation.bias"]

    s = {name: style(name, li) for name, _, li, _ in convs}
    s.update({name: style(name, li) for name, li in rgbs})

    x = sd["input.input"][0].permute(1, 2, 0)[None] * s["conv1"][:, None, None, :]  # x~0 NHWC
    acts, skips = {}, {}
    skip, rgb_i = None, 0
    for idx, (name, up, _, ni) in enumerate(convs):
        w = sd[name + ".conv.weight"][0]
        cout, cin = w.shape[0], w.shape[1]
        wt = w * (1 / math.sqrt(cin * 9))
        wsq = (wt * wt).sum([2, 3])
        d = torch.rsqrt((s[name] ** 2) @ wsq.t() + 1e-8)                      # [B, Cout]
        nxt = convs[idx + 1][0] if idx + 1 < len(convs) else None
        nz = noise[ni].to(dtype)
        nw, bias = sd[name + ".noise.weight"], sd[name + ".activate.bias"]
        H = x.shape[1]
        if up and composite:
            if composite == "row":
                v = _rowfold_upconv(x, rowfold_weights(wt, f), f, H) * d[:, None, None, :]
            else:
                v = _composite_upconv(x, composite_weights(wt, f), H) * d[:, None, None, :]
            v = v + nw * nz.permute(0, 2, 3, 1) + bias
            y = torch.where(v > 0, v, 0.2 * v) * math.sqrt(2)
            acts[name] = y.permute(0, 3, 1, 2)
            x = y * s[nxt][:, None, None, :]
            continue
        if up:
            t = x.new_zeros(B, 2 * H + 2, 2 * H + 2, cout)
            for ph in range(4):
                py, px = ph >> 1, ph & 1
                acc = _tap_conv(x, wt, _taps_up(py, px), H + 1, H + 1)
                t[:, py::2, px::2] = acc * d[:, None, None, :]
            OH = 2 * H
            v = x.new_zeros(B, OH, OH, cout)
            for i in range(4):
                for j in range(4):
                    # out[Y][X] += f[i] f[j] t[Y+i-1][X+j-1]
                    ys, xs = max(0, 1 - i), max(0, 1 - j)
                    src = t[:, ys + i - 1:OH + i - 1, xs + j - 1:OH + j - 1]
                    v[:, ys:ys + src.shape[1], xs:xs + src.shape[2]] += f[i] * f[j] * src
            v = v + nw * nz.permute(0, 2, 3, 1) + bias
            y = torch.where(v > 0, v, 0.2 * v) * math.sqrt(2)
            acts[name] = y.permute(0, 3, 1, 2)
            x = y * s[nxt][:, None, None, :]
            continue
        acc = _tap_conv(x, wt, _taps_plain(), H, H)
        v = acc * d[:, None, None, :] + nw * nz.permute(0, 2, 3, 1) + bias
        y = torch.where(v > 0, v, 0.2 * v) * math.sqrt(2)
        acts[name] = y.permute(0, 3, 1, 2)
        rname, _ = rgbs[rgb_i]
        wr = sd[rname + ".conv.weight"][0, :, :, 0, 0] * (1 / math.sqrt(cout))     # [3, C]
        wr_b = wr[None] * s[rname][:, None, :]                                       # [B, 3, C]
        rgb = torch.einsum("bhwc,bkc->bkhw", y, wr_b) + sd[rname + ".bias"]
        if skip is not None:
            rgb = rgb + _upsample2x(skip, f)
        skip = rgb
        skips[rgb_i] = skip
        rgb_i += 1
        if nxt is not None:
            x = y * s[nxt][:, None, None, :]
    return skip, acts, skips
