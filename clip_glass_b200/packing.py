"""Host-side weight packing: reference state_dict layouts -> engine tensors.

Runs once at ``Generator.__init__`` time (it replaces ``stylegan2.models.load``
+ ``clip.load`` of generator.py:16-19), on the CPU, in fp32; nothing here is on
the per-generation path.  Every packed tensor is uploaded with
``glass_set_tensor(name, ...)``; names and layouts below are the contract with
clip_glass_b200/csrc/engine.cu.

Algebra used (derivations in DESIGN.md §3):

* Modulated conv (stylegan2/modules.py:920-967).  The reference builds a
  per-sample weight  W'[b,o,i,k] = c W[o,i,k] s[b,i] d[b,o].  Here the
  weights stay batch-shared: the *input* is pre-scaled by s[b,i] (done by the
  producing kernel's epilogue), the accumulator is scaled by
  d[b,o] = rsqrt(sum_i s[b,i]^2 Wsq[o,i] + eps),  Wsq[o,i] = sum_k (c W[o,i,k])^2.
* Up-conv (modules.py:1089-1139): conv_transpose2d(stride 2, 3x3) followed by
  the 4x4 FIR (pad 1) equals ONE 3x3 conv over the un-upsampled input with
  4*Cout output channels (one group per output phase (py,px)) followed by a
  depth-to-space:   out[2z+p] = sum_{a in 0..2} g[p + 2 - 2a] x[z + a - 1],
  g[t] = sum_{k-j+1=t} w[k] f[j]  (t in -2..3), per axis.
* D down-conv (modules.py:1238-1254): FIR (pad 2) then 3x3 stride-2 conv equals
  ONE 3x3 conv over the space-to-depth(2) input (4*Cin channels):
  out[z] = sum_{a,p} g[2a + p] X[z + a - 1, p],  g[t] = sum_{k+j=t} w[k] f[j].
"""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np
import torch

from .weights import ClipSpec, GanSpec

F1_UP = torch.tensor([1.0, 3.0, 3.0, 1.0]) / 4.0     # per-axis FIR, gain*up^2 = 4 (modules.py:1050-1055)
F1_DOWN = torch.tensor([1.0, 3.0, 3.0, 1.0]) / 8.0   # per-axis FIR, gain 1 (modules.py:1198-1203)


def _coef(shape, lr_mul=1.0):
    return lr_mul / math.sqrt(float(np.prod(shape[1:])))


def g_layers(spec: GanSpec):
    """Forward-order list of the modulated 3x3 convs of G:
    dicts(block, idx_in_block, cin, cout, up, res_out)."""
    ch = list(spec.channels)[::-1]
    out = []
    for b in range(spec.num_blocks):
        res = 4 * 2 ** b
        if b == 0:
            out.append(dict(block=0, l=0, cin=ch[0], cout=ch[0], up=False, res=res))
        else:
            out.append(dict(block=b, l=0, cin=ch[b - 1], cout=ch[b], up=True, res=res))
            out.append(dict(block=b, l=1, cin=ch[b], cout=ch[b], up=False, res=res))
    return out


def style_offsets(spec: GanSpec):
    """Column offsets into the concatenated style vector: conv layers in
    forward order, then the toRGB layers by block.  Returns (conv_off, rgb_off, total)."""
    ch = list(spec.channels)[::-1]
    off, conv_off, rgb_off = 0, [], []
    for ly in g_layers(spec):
        conv_off.append(off)
        off += ly["cin"]
    for b in range(spec.num_blocks):
        rgb_off.append(off)
        off += ch[b]
    return conv_off, rgb_off, off


def fold_upconv(w: torch.Tensor) -> torch.Tensor:
    """w [O,I,3,3] (already * coef) -> [9 taps][4*O][I], tap = a*3+b,
    n = (py*2+px)*O + o."""
    O, I = w.shape[:2]
    # G2[ty+2][tx+2] = sum_{ky-jy+1=ty, kx-jx+1=tx} w[ky,kx] f[jy] f[jx]
    G2 = torch.zeros(O, I, 6, 6, dtype=w.dtype)
    for ky in range(3):
        for jy in range(4):
            ty = ky - jy + 1
            for kx in range(3):
                for jx in range(4):
                    tx = kx - jx + 1
                    G2[:, :, ty + 2, tx + 2] += w[:, :, ky, kx] * (F1_UP[jy] * F1_UP[jx])
    out = torch.zeros(9, 4 * O, I, dtype=w.dtype)
    for a in range(3):
        for b in range(3):
            for py in range(2):
                for px in range(2):
                    ph = py * 2 + px
                    out[a * 3 + b, ph * O:(ph + 1) * O] = G2[:, :, py + 4 - 2 * a, px + 4 - 2 * b]
    return out


def fold_downconv(w: torch.Tensor) -> torch.Tensor:
    """w [O,I,3,3] (already * coef) -> [9 taps][O][4*I], k = (py*2+px)*I + i."""
    O, I = w.shape[:2]
    G2 = torch.zeros(O, I, 6, 6, dtype=w.dtype)
    for ky in range(3):
        for jy in range(4):
            for kx in range(3):
                for jx in range(4):
                    G2[:, :, ky + jy, kx + jx] += w[:, :, ky, kx] * (F1_DOWN[jy] * F1_DOWN[jx])
    out = torch.zeros(9, O, 4 * I, dtype=w.dtype)
    for a in range(3):
        for b in range(3):
            for py in range(2):
                for px in range(2):
                    ph = py * 2 + px
                    out[a * 3 + b, :, ph * I:(ph + 1) * I] = G2[:, :, 2 * a + py, 2 * b + px]
    return out


# Tap tables of the exact polyphase forms (dy, dx per tap); see DESIGN.md section 3.
UP_EXACT_TAPS = [(-1, -1), (-1, 0), (0, -1), (0, 0)]
DOWN_EXACT_TAPS = [(0, 0), (0, 1), (1, 0), (1, 1)]


def exact_upconv(w: torch.Tensor) -> torch.Tensor:
    """conv_transpose2d(stride 2, 3x3) WITHOUT the FIR, as a 2x2-tap conv with 4*O phase columns.

    u[2z+py, 2w+px] = sum over (y,ky): 2y+ky = 2z+py  ->  y = z (ky = py)  or  y = z-1 (ky = 2, only py = 0);
    same along x.  w [O,I,3,3] (already * coef) -> [4 taps][4*O][I], tap order UP_EXACT_TAPS,
    column (py*2+px)*O + o.  9 of the 16 (tap, phase) blocks are non-zero: the 9 MACs/pixel of the
    transposed conv as the reference writes it."""
    O, I = w.shape[:2]
    out = torch.zeros(4, 4 * O, I, dtype=w.dtype)
    for t, (dy, dx) in enumerate(UP_EXACT_TAPS):
        for py in range(2):
            ky = py if dy == 0 else (2 if py == 0 else None)
            if ky is None:
                continue
            for px in range(2):
                kx = px if dx == 0 else (2 if px == 0 else None)
                if kx is None:
                    continue
                ph = py * 2 + px
                out[t, ph * O:(ph + 1) * O] = w[:, :, ky, kx]
    return out


def exact_downconv(w: torch.Tensor) -> torch.Tensor:
    """3x3 stride-2 conv (no FIR) over the space-to-depth(2) of the blurred input: out[z] = sum_k w[k] u[2z+k]
    = sum_{a,p: 2a+p=k} w[2a+p] U[z+a, p].  w [O,I,3,3] -> [4 taps][O][4*I], tap order DOWN_EXACT_TAPS,
    k index (py*2+px)*I + i."""
    O, I = w.shape[:2]
    out = torch.zeros(4, O, 4 * I, dtype=w.dtype)
    for t, (a, b) in enumerate(DOWN_EXACT_TAPS):
        for py in range(2):
            ky = 2 * a + py
            if ky > 2:
                continue
            for px in range(2):
                kx = 2 * b + px
                if kx > 2:
                    continue
                ph = py * 2 + px
                out[t, :, ph * I:(ph + 1) * I] = w[:, :, ky, kx]
    return out


def pair_pack(w: torch.Tensor) -> torch.Tensor:
    """3x3 conv over horizontally PAIRED pixels: the NHWC tensor [H][W][C] is read as [H][W/2][2C] and written as
    [H][W/2][2*O] (the same bytes).  Row xp of the GEMM holds pixels 2xp, 2xp+1; tap (ky, kxp) reads pair xp+kxp-1.
    w [O,I,3,3] (already * coef) -> [9 taps][2*O][2*I] with
        W[(ky,kxp)][px_out*O + o][px_in*I + i] = w[o,i,ky,kx],  kx = 2*(kxp-1) + px_in - px_out + 1  (if 0 <= kx <= 2).
    Half of the blocks are zero (2x the MACs), in exchange for 128-byte TMA rows and 256-pixel tiles on the
    32-channel layers, which are bound by TMA row rate and per-tile bookkeeping, not by math."""
    O, I = w.shape[:2]
    out = torch.zeros(9, 2 * O, 2 * I, dtype=w.dtype)
    for ky in range(3):
        for kxp in range(3):
            for po in range(2):
                for pi in range(2):
                    kx = 2 * (kxp - 1) + pi - po + 1
                    if 0 <= kx <= 2:
                        out[ky * 3 + kxp, po * O:(po + 1) * O, pi * I:(pi + 1) * I] = w[:, :, ky, kx]
    return out


def taps_plain(w: torch.Tensor) -> torch.Tensor:
    """w [O,I,k,k] -> [k*k][O][I]."""
    O, I, k, _ = w.shape
    return w.permute(2, 3, 0, 1).reshape(k * k, O, I).contiguous()


def _f32(t):
    return np.ascontiguousarray(t.detach().float().numpy())


def _f16(t):
    return np.ascontiguousarray(t.detach().float().half().numpy())


def pack_generator(sd: Dict[str, torch.Tensor], spec: GanSpec) -> Dict[str, np.ndarray]:
    out: Dict[str, np.ndarray] = {}
    L = spec.latent_size
    for i in range(spec.mapping_layers):
        w = sd[f"G_mapping.main.{i}.layer.weight"].float()
        out[f"g.map.w{i}"] = _f32((w * _coef(w.shape, 0.01)).t())          # [in][out]
        out[f"g.map.b{i}"] = _f32(sd[f"G_mapping.main.{i}.bias"].float() * 0.01)
    conv_off, rgb_off, S = style_offsets(spec)
    style_w = torch.zeros(L, S)
    style_b = torch.zeros(S)
    layers = g_layers(spec)
    for li, ly in enumerate(layers):
        p = f"G_synthesis.conv_blocks.{ly['block']}.conv_block.{ly['l']}"
        a = sd[p + ".layer.layer.dense.layer.weight"].float()              # [cin][L]
        style_w[:, conv_off[li]:conv_off[li] + ly["cin"]] = (a * _coef(a.shape)).t()
        style_b[conv_off[li]:conv_off[li] + ly["cin"]] = sd[p + ".layer.layer.dense.bias"].float()
        w = sd[p + ".layer.layer.weight"].float()
        wc = w * _coef(w.shape)
        out[f"g.conv{li}.wsq"] = _f32((wc ** 2).sum(dim=(2, 3)).t())       # [cin][cout]
        out[f"g.conv{li}.w"] = _f16(fold_upconv(wc) if ly["up"] else taps_plain(wc))
        if ly["up"]:
            out[f"g.conv{li}.wx"] = _f16(exact_upconv(wc))      # exact polyphase form (used from 16x16 inputs up)
        elif ly["cin"] == 32 and ly["res"] >= 32:
            out[f"g.conv{li}.wp"] = _f16(pair_pack(wc))         # pixel-pair form of the 32-channel layers
        out[f"g.conv{li}.bias"] = _f32(sd[p + ".bias"])
        out[f"g.conv{li}.nstr"] = _f32(sd[p + ".layer.weight"].reshape(1))
    ch = list(spec.channels)[::-1]
    for b in range(spec.num_blocks):
        p = f"G_synthesis.to_data_layers.{b}"
        a = sd[p + ".layer.dense.layer.weight"].float()
        style_w[:, rgb_off[b]:rgb_off[b] + ch[b]] = (a * _coef(a.shape)).t()
        style_b[rgb_off[b]:rgb_off[b] + ch[b]] = sd[p + ".layer.dense.bias"].float()
        w = sd[p + ".layer.weight"].float()
        out[f"g.rgb{b}.w"] = _f32((w * _coef(w.shape)).reshape(3, ch[b]))
        out[f"g.rgb{b}.bias"] = _f32(sd[p + ".bias"])
    out["g.style.w"] = _f32(style_w)
    out["g.style.b"] = _f32(style_b)
    out["g.const"] = _f32(sd["G_synthesis.const"].float().permute(1, 2, 0).reshape(16, ch[0]))
    return out


def pad_to(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def pack_discriminator(sd: Dict[str, torch.Tensor], spec: GanSpec) -> Dict[str, np.ndarray]:
    out: Dict[str, np.ndarray] = {}
    ch = list(spec.channels)
    nb = spec.num_blocks
    w = sd["from_data_layers.0.layer.weight"].float()
    out["d.frgb.w"] = _f32((w * _coef(w.shape)).reshape(ch[0], 3).t())       # [3][C0]
    out["d.frgb.b"] = _f32(sd["from_data_layers.0.bias"])
    for b in range(nb - 1):
        p = f"conv_blocks.{b}"
        w0 = sd[p + ".conv_block.0.layer.weight"].float()
        out[f"d.b{b}.c0.w"] = _f16(taps_plain(w0 * _coef(w0.shape)))
        if ch[b] == 32 and (spec.resolution >> b) >= 32:
            out[f"d.b{b}.c0.wp"] = _f16(pair_pack(w0 * _coef(w0.shape)))
        out[f"d.b{b}.c0.b"] = _f32(sd[p + ".conv_block.0.bias"])
        w1 = sd[p + ".conv_block.1.layer.weight"].float()
        out[f"d.b{b}.c1.w"] = _f16(fold_downconv(w1 * _coef(w1.shape)))
        out[f"d.b{b}.c1.wx"] = _f16(exact_downconv(w1 * _coef(w1.shape)))
        if ch[b] in (32, 64):
            out[f"d.b{b}.c1.w9"] = _f16(taps_plain(w1 * _coef(w1.shape)))     # fused-FIR exact down-conv (downconv_tc.cu)
        out[f"d.b{b}.c1.b"] = _f32(sd[p + ".conv_block.1.bias"])
        wp = sd[p + ".projection.weight"].float()
        out[f"d.b{b}.proj.w"] = _f16(taps_plain(wp * _coef(wp.shape)))
    p = f"conv_blocks.{nb - 1}.1.conv_block.0"
    w = sd[p + ".layer.weight"].float()
    wc = taps_plain(w * _coef(w.shape))                                       # [9][C][C+1]
    C = ch[-1]
    cpad = pad_to(w.shape[1], 64)
    wpad = torch.zeros(9, C, cpad)
    wpad[:, :, :w.shape[1]] = wc
    out["d.fin.w"] = _f16(wpad)
    out["d.fin.b"] = _f32(sd[p + ".bias"])
    w = sd["dense.0.layer.weight"].float()                                    # [C][C*16], cols c*16+y*4+x
    wd = (w * _coef(w.shape)).reshape(C, C, 16).permute(0, 2, 1).reshape(1, C, 16 * C)
    out["d.dense0.w"] = _f16(wd)                                              # cols (y*4+x)*C + c
    out["d.dense0.b"] = _f32(sd["dense.0.bias"])
    w = sd["dense.1.layer.weight"].float()
    out["d.dense1.w"] = _f32((w * _coef(w.shape)).reshape(C))
    out["d.dense1.b"] = _f32(sd["dense.1.bias"].reshape(1))
    return out


def pack_clip_visual(sd: Dict[str, torch.Tensor], spec: ClipSpec) -> Dict[str, np.ndarray]:
    """``sd`` is the visual tower state dict (fp32 master or as-built); GEMM
    weights are rounded to fp16 exactly as convert_weights does
    (clip/model.py:339-360); LayerNorm / cls / pos stay fp32."""
    out: Dict[str, np.ndarray] = {}
    Wd = spec.width
    h = lambda t: t.float().half()
    out["c.patch.w"] = _f16(h(sd["conv1.weight"]).reshape(1, Wd, -1))
    out["c.cls"] = _f32(sd["class_embedding"])
    out["c.pos"] = _f32(sd["positional_embedding"])
    out["c.lnpre.w"] = _f32(sd["ln_pre.weight"])
    out["c.lnpre.b"] = _f32(sd["ln_pre.bias"])
    for l in range(spec.layers):
        p = f"transformer.resblocks.{l}"
        out[f"c.l{l}.ln1.w"] = _f32(sd[p + ".ln_1.weight"])
        out[f"c.l{l}.ln1.b"] = _f32(sd[p + ".ln_1.bias"])
        out[f"c.l{l}.qkv.w"] = _f16(h(sd[p + ".attn.in_proj_weight"])[None])
        out[f"c.l{l}.qkv.b"] = _f32(h(sd[p + ".attn.in_proj_bias"]))
        out[f"c.l{l}.out.w"] = _f16(h(sd[p + ".attn.out_proj.weight"])[None])
        out[f"c.l{l}.out.b"] = _f32(h(sd[p + ".attn.out_proj.bias"]))
        out[f"c.l{l}.ln2.w"] = _f32(sd[p + ".ln_2.weight"])
        out[f"c.l{l}.ln2.b"] = _f32(sd[p + ".ln_2.bias"])
        out[f"c.l{l}.fc.w"] = _f16(h(sd[p + ".mlp.c_fc.weight"])[None])
        out[f"c.l{l}.fc.b"] = _f32(h(sd[p + ".mlp.c_fc.bias"]))
        out[f"c.l{l}.proj.w"] = _f16(h(sd[p + ".mlp.c_proj.weight"])[None])
        out[f"c.l{l}.proj.b"] = _f32(h(sd[p + ".mlp.c_proj.bias"]))
    out["c.lnpost.w"] = _f32(sd["ln_post.weight"])
    out["c.lnpost.b"] = _f32(sd["ln_post.bias"])
    out["c.proj"] = _f32(h(sd["proj"]))                                       # [W][E]
    return out
