"""CPU emulation of the ENGINE'S algorithm (not of the reference's): the same
packed tensors, the same re-associations (pre-scaled inputs + demod on the
accumulator, phase-folded up-conv, space-to-depth down-conv, fused toRGB
partials, polyphase skip upsample) and the same fp16 rounding points as the
CUDA kernels in clip_glass_b200/csrc, written with torch ops.

Purpose: (1) prove on the CPU, against the oracle, that the packing algebra in
clip_glass_b200/packing.py is exact; (2) measure how far fp16 storage moves
the final scores, i.e. what tolerance the CUDA path can meet; (3) give the GPU
tests a layer-by-layer expectation with identical rounding to localise a
kernel bug.  Test infrastructure only.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from clip_glass_b200 import packing

SQRT2 = math.sqrt(2.0)


def T(a):
    return torch.from_numpy(np.asarray(a))


def h(x):
    """round to fp16 storage, continue in fp32"""
    return x.half().float()


def conv_taps(x_nhwc: torch.Tensor, w_taps: torch.Tensor) -> torch.Tensor:
    """x [N,H,W,C] (fp32 holding fp16 values), w [taps][Ntot][C] -> fp32 [N,H,W,Ntot];
    tap = ky*3+kx reads input offset (ky-1, kx-1), zero outside."""
    taps, ntot, c = w_taps.shape
    k = int(round(math.sqrt(taps)))
    w = w_taps.float().reshape(k, k, ntot, c).permute(2, 3, 0, 1)
    y = F.conv2d(x_nhwc.permute(0, 3, 1, 2), w, padding=k // 2)
    return y.permute(0, 2, 3, 1)


def conv_taps_table(x_nhwc: torch.Tensor, w_taps: torch.Tensor, table, out_hw) -> torch.Tensor:
    """Generic tap-table conv: out[n,y,x,:] = sum_t W[t] @ x[n, y+dy_t, x+dx_t, :] (zero outside the input).
    x [N,Hin,Win,C]; out grid out_hw = (H, W)."""
    n, hin, win, c = x_nhwc.shape
    H, Wd = out_hw
    ntot = w_taps.shape[1]
    out = torch.zeros(n, H, Wd, ntot)
    pad = 2
    xp = F.pad(x_nhwc, [0, 0, pad, pad + max(0, Wd - win) + 2, pad, pad + max(0, H - hin) + 2])
    for t, (dy, dx) in enumerate(table):
        patch = xp[:, pad + dy:pad + dy + H, pad + dx:pad + dx + Wd, :]
        out += patch @ w_taps[t].float().t()
    return out


def fir_same(u: torch.Tensor, f1: torch.Tensor, out_hw, shift: int) -> torch.Tensor:
    """v[Z,X] = sum_{jy,jx} f[jy] f[jx] u[Z+jy-shift, X+jx-shift] (zero outside u). u [N,Hu,Wu,C]."""
    n, hu, wu, c = u.shape
    H, Wd = out_hw
    k = (f1[:, None] * f1[None, :])
    up = F.pad(u.permute(0, 3, 1, 2), [shift, 4, shift, 4])
    v = F.conv2d(up, k[None, None].repeat(c, 1, 1, 1), groups=c)
    return v[:, :, :H, :Wd].permute(0, 2, 3, 1)


def depth_to_space(y: torch.Tensor, cout: int) -> torch.Tensor:
    """[N,H,W,4*C] with n=(py*2+px)*C+o -> [N,2H,2W,C]"""
    n, hh, ww, _ = y.shape
    y = y.reshape(n, hh, ww, 2, 2, cout).permute(0, 1, 3, 2, 4, 5)
    return y.reshape(n, 2 * hh, 2 * ww, cout)


def space_to_depth(x: torch.Tensor) -> torch.Tensor:
    """[N,H,W,C] -> [N,H/2,W/2,4C] with k=(py*2+px)*C+i"""
    n, hh, ww, c = x.shape
    x = x.reshape(n, hh // 2, 2, ww // 2, 2, c).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(n, hh // 2, ww // 2, 4 * c)


def skip_upsample(y: torch.Tensor) -> torch.Tensor:
    """[N,H,W,C] -> [N,2H,2W,C]: v[2z] = .75 x[z-1] + .25 x[z]; v[2z+1] = .25 x[z-1] + .75 x[z]
    per axis with x[-1] = 0 (derived from modules.py:569-602)."""
    def up1(t, dim):
        prev = torch.roll(t, 1, dims=dim)
        idx = [slice(None)] * t.dim()
        idx[dim] = 0
        prev[tuple(idx)] = 0
        even = 0.75 * prev + 0.25 * t
        odd = 0.25 * prev + 0.75 * t
        st = torch.stack([even, odd], dim=dim + 1)
        shp = list(t.shape)
        shp[dim] *= 2
        return st.reshape(shp)
    return up1(up1(y, 1), 2)


def lrelu_gain(x):
    return F.leaky_relu(x, 0.2) * SQRT2


def emu_mapping(pk, spec, z):
    x = z * torch.rsqrt((z * z).mean(-1, keepdim=True) + 1e-8)
    for i in range(spec.mapping_layers):
        x = lrelu_gain(x @ T(pk[f"g.map.w{i}"]) + T(pk[f"g.map.b{i}"]))
    return x


def emu_generator(pk, spec, z, noise, batch, capture=None, fp16=True, exact=True):
    """z [P,L] fp32; noise list per group of list per layer [1,1,H,W] or None.
    Returns images [P,3,R,R] in [0,1]."""
    r = h if fp16 else (lambda t: t)
    P = z.shape[0]
    w = emu_mapping(pk, spec, z)
    styles = w @ T(pk["g.style.w"]) + T(pk["g.style.b"])                  # [P,S]
    conv_off, rgb_off, _ = packing.style_offsets(spec)
    layers = packing.g_layers(spec)
    ch = list(spec.channels)[::-1]
    if capture is not None:
        capture["w"] = w
        capture["styles"] = styles
    def pow2_scale(sl):
        """k_style_norm: m = the power of two with max|s| / m in [0.5, 1) per sample (1 for an all-zero slice)."""
        mx = sl.abs().amax(dim=1, keepdim=True)
        _, ex = torch.frexp(mx)
        return torch.where(mx > 0, torch.ldexp(torch.ones_like(mx), ex), torch.ones_like(mx))

    s0 = styles[:, conv_off[0]:conv_off[0] + ch[0]]
    s0 = s0 / pow2_scale(s0)
    x = r(T(pk["g.const"]).reshape(1, 4, 4, ch[0]) * s0[:, None, None, :])   # pre-scaled input
    if capture is not None:
        capture["x0"] = x
    y = None
    for li, ly in enumerate(layers):
        cin, cout = ly["cin"], ly["cout"]
        s = styles[:, conv_off[li]:conv_off[li] + cin]
        d = torch.rsqrt((s * s) @ T(pk[f"g.conv{li}.wsq"]) + 1e-8) * pow2_scale(s)   # [P,cout]; x carries s / m
        in_res = x.shape[1]
        if ly["up"] and exact and in_res >= 16:
            # exact polyphase transposed conv on the (H+1)x(W+1) grid -> u (fp16), then the FIR pass
            acc = conv_taps_table(x, T(pk[f"g.conv{li}.wx"]), packing.UP_EXACT_TAPS, (in_res + 1, in_res + 1))
            u = r(depth_to_space(acc * d[:, None, None, :].repeat(1, 1, 1, 4), cout))      # [(2H+2)^2], fp16 storage
            if capture is not None:
                capture[f"u{li}"] = u
            v = fir_same(u, packing.F1_UP, (2 * in_res, 2 * in_res), 1)
        else:
            acc = conv_taps(x, T(pk[f"g.conv{li}.w"]))
            if ly["up"]:
                acc = depth_to_space(acc, cout)
            v = acc * d[:, None, None, :]
        if noise is not None:
            nz = torch.cat([noise[g][li].reshape(1, ly["res"], ly["res"], 1).expand(batch, -1, -1, -1)
                            for g in range(P // batch)])
            v = v + T(pk[f"g.conv{li}.nstr"]) * nz
        v = lrelu_gain(v + T(pk[f"g.conv{li}.bias"]))
        if capture is not None:
            capture[f"act{li}"] = v
        last_in_block = (li + 1 == len(layers)) or (layers[li + 1]["block"] != ly["block"])
        if last_in_block:
            b = ly["block"]
            srgb = styles[:, rgb_off[b]:rgb_off[b] + cout]
            wr = T(pk[f"g.rgb{b}.w"])[None] * srgb[:, None, :]               # [P,3,C]
            t = torch.einsum("nhwc,nkc->nhwk", v, wr) + T(pk[f"g.rgb{b}.bias"])
            y = t if y is None else skip_upsample(y) + t
            if capture is not None:
                capture[f"rgb{b}"] = y
        if li + 1 < len(layers):
            sn = styles[:, conv_off[li + 1]:conv_off[li + 1] + cout]
            sn = sn / pow2_scale(sn)
            x = r(v * sn[:, None, None, :])
            if capture is not None:
                capture[f"xs{li}"] = x
    img = ((y + 1) / 2).clip(0, 1)
    return img.permute(0, 3, 1, 2).contiguous()


def emu_discriminator(pk, spec, images, batch, capture=None, fp16=True, exact=True):
    """images [P,3,R,R] in [0,1] -> logits [P]."""
    r = h if fp16 else (lambda t: t)
    ch = list(spec.channels)
    P = images.shape[0]
    den = images * 2 - 1
    x = torch.einsum("nchw,ck->nhwk", den, T(pk["d.frgb.w"])) + T(pk["d.frgb.b"])
    x = r(lrelu_gain(x))
    f1 = packing.F1_DOWN
    blur = (f1[:, None] * f1[None, :])
    for b in range(spec.num_blocks - 1):
        a = r(lrelu_gain(conv_taps(x, T(pk[f"d.b{b}.c0.w"])) + T(pk[f"d.b{b}.c0.b"])))
        res = a.shape[1]
        use_exact = exact and res >= 32
        # projection path: FIR(pad 1) sampled at even positions, then 1x1
        c = x.shape[-1]
        xp = F.pad(x.permute(0, 3, 1, 2), [1, 1, 1, 1])
        xd = F.conv2d(xp, blur[None, None].repeat(c, 1, 1, 1), stride=2, groups=c)
        xd = r(xd.permute(0, 2, 3, 1))
        proj = r(conv_taps(xd, T(pk[f"d.b{b}.proj.w"])))
        if use_exact:
            # blur (FIR pad 2) -> (H+1)^2, stored space-to-depth as [(H/2+1)^2][4C] fp16; then the 2x2-tap conv
            u = fir_same(a, packing.F1_DOWN, (res + 1, res + 1), 2)
            up = F.pad(u, [0, 0, 0, 1, 0, 1])                                    # pad to (H+2)^2 (zeros)
            u_s2d = r(space_to_depth(up))
            acc = conv_taps_table(u_s2d, T(pk[f"d.b{b}.c1.wx"]), packing.DOWN_EXACT_TAPS, (res // 2, res // 2))
        else:
            acc = conv_taps(space_to_depth(a), T(pk[f"d.b{b}.c1.w"]))
        x = r((lrelu_gain(acc + T(pk[f"d.b{b}.c1.b"])) + proj) * (1.0 / SQRT2))
        if capture is not None:
            capture[f"d{b}"] = x
    # minibatch std incl. the reference's in-place centring quirk
    G = spec.mbstd_group_size
    C = ch[-1]
    xs = x.reshape(P // batch, G, batch // G, 4, 4, C)
    cen = xs - xs.mean(dim=1, keepdim=True)
    std = torch.sqrt((cen ** 2).mean(dim=1) + 1e-8).mean(dim=(2, 3, 4))       # [groups, batch//G]
    feat = std[:, None, :, None, None, None].expand(-1, G, -1, 4, 4, 1)
    cpad = pk["d.fin.w"].shape[-1]
    xin = torch.zeros(P, 4, 4, cpad)
    xin[..., :C] = cen.reshape(P, 4, 4, C)
    xin[..., C:C + 1] = feat.reshape(P, 4, 4, 1)
    xin = r(xin)
    x = r(lrelu_gain(conv_taps(xin, T(pk["d.fin.w"])) + T(pk["d.fin.b"])))
    x = x.reshape(P, 1, 1, 16 * C)
    x = r(lrelu_gain(conv_taps(x, T(pk["d.dense0.w"])) + T(pk["d.dense0.b"]))).reshape(P, C)
    return x @ T(pk["d.dense1.w"]) + T(pk["d.dense1.b"])


def emu_resize_patches(images, spec):
    """[P,3,R,R] -> patch matrix [P*49, 3*p*p] fp16-valued; bilinear align_corners=False."""
    P = images.shape[0]
    R = spec.resolution
    p = spec.patch
    g = R // p
    img = h(F.interpolate(images, size=(R, R), mode="bilinear", align_corners=False))
    pt = img.reshape(P, 3, g, p, g, p).permute(0, 2, 4, 1, 3, 5).reshape(P * g * g, 3 * p * p)
    return pt


def emu_clip(pk, spec, images, text, capture=None):
    """returns (features [P,E] fp32 (fp16-valued), sim [P] fp32)."""
    P = images.shape[0]
    Wd, Tn, H = spec.width, spec.tokens, spec.heads
    pt = emu_resize_patches(images, spec)
    emb = h(pt @ T(pk["c.patch.w"])[0].float().t()).reshape(P, Tn - 1, Wd)
    cls = h(T(pk["c.cls"])).reshape(1, 1, Wd).expand(P, 1, Wd)
    x = torch.cat([cls, emb], dim=1)
    x = h(x + h(T(pk["c.pos"])))
    ln = lambda t, w, b: h(F.layer_norm(t, (Wd,), T(pk[w]), T(pk[b]), 1e-5))
    x = ln(x, "c.lnpre.w", "c.lnpre.b")
    if capture is not None:
        capture["ln_pre"] = x
    for l in range(spec.layers):
        q = f"c.l{l}."
        hh = ln(x, q + "ln1.w", q + "ln1.b")
        qkv = h(hh @ T(pk[q + "qkv.w"])[0].float().t() + T(pk[q + "qkv.b"]))
        qq, kk, vv = qkv.reshape(P, Tn, 3, H, 64).permute(2, 0, 3, 1, 4)
        att = torch.softmax((qq * 0.125) @ kk.transpose(-1, -2), dim=-1)
        o = h((att @ vv).permute(0, 2, 1, 3).reshape(P, Tn, Wd))
        o = h(o @ T(pk[q + "out.w"])[0].float().t() + T(pk[q + "out.b"]))
        x = h(x + o)
        hh = ln(x, q + "ln2.w", q + "ln2.b")
        f = h(hh @ T(pk[q + "fc.w"])[0].float().t() + T(pk[q + "fc.b"]))
        f = h(f * torch.sigmoid(1.702 * f))
        o = h(f @ T(pk[q + "proj.w"])[0].float().t() + T(pk[q + "proj.b"]))
        x = h(x + o)
        if capture is not None:
            capture[f"block{l}"] = x
    c = ln(x[:, 0], "c.lnpost.w", "c.lnpost.b")
    feats = h(c @ T(pk["c.proj"]))
    t = text.float().reshape(1, -1)
    sim = (feats * t).sum(-1) / torch.clamp(feats.norm(dim=-1) * t.norm(dim=-1), min=1e-8)
    return feats, sim
