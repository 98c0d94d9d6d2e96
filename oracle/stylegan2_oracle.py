"""ORACLE (test infrastructure only — never on the product path).

CPU fp32 restatement of the StyleGAN2 arithmetic that sits on CLIP-GLaSS's
fitness path.  Written from scratch as plain functions over the reference's
state_dict (same key names), each citing the reference lines it follows.
Pinned against the reference's own modules by ``oracle/make_golden.py``
(which imports /root/reference in the build container and commits the
outputs under tests/golden/), and re-checked against those fixtures by
``tests/test_oracle.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2.0)
EPS = 1e-8


# ----------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------
def _coef(shape: Sequence[int], lr_mul: float = 1.0, gain: float = 1.0) -> float:
    """Equalised-learning-rate runtime coefficient.
    stylegan2/modules.py:103-108 (weight_scale=True): he_std * lr_mul."""
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    return gain / math.sqrt(fan_in) * lr_mul


def fir_kernel(gain: float = 1.0, up_factor: int = 1) -> torch.Tensor:
    """[1,3,3,1] (x) [1,3,3,1], normalised, times gain*up^2.
    stylegan2/modules.py:169-203."""
    f = torch.tensor([1.0, 3.0, 3.0, 1.0])
    k = f[:, None] * f[None, :]
    k = k / k.sum()
    return k * (gain * up_factor ** 2)


def fir(x: torch.Tensor, kernel: torch.Tensor, pad0: int, pad1: int, stride: int = 1) -> torch.Tensor:
    """Depthwise FIR.  stylegan2/modules.py:499-523 (FilterLayer.forward)."""
    c = x.shape[1]
    x = F.pad(x, [pad0, pad1, pad0, pad1])
    return F.conv2d(x, kernel[None, None].repeat(c, 1, 1, 1), stride=stride, groups=c)


def lrelu_gain(x: torch.Tensor) -> torch.Tensor:
    """leaky 0.2 then *sqrt(2).  stylegan2/modules.py:24-31, 294-296."""
    return F.leaky_relu(x, 0.2) * SQRT2


# ----------------------------------------------------------------------------
# mapping network
# ----------------------------------------------------------------------------
def mapping(z: torch.Tensor, sd: Dict[str, torch.Tensor], num_layers: int = 8) -> torch.Tensor:
    """stylegan2/models.py:590-627: pixel-norm, then num_layers x
    (dense * 0.01/sqrt(fan_in); + 0.01*bias; lrelu 0.2; *sqrt2)."""
    x = z * torch.rsqrt(torch.mean(z * z, dim=-1, keepdim=True) + EPS)
    for i in range(num_layers):
        w = sd[f"G_mapping.main.{i}.layer.weight"]
        b = sd[f"G_mapping.main.{i}.bias"]
        x = x.matmul((w * _coef(w.shape, lr_mul=0.01)).t())     # modules.py:795-798
        x = x + 0.01 * b                                         # modules.py:289-293 (bias_coef = lr_mul)
        x = lrelu_gain(x)
    return x


# ----------------------------------------------------------------------------
# modulated convolution
# ----------------------------------------------------------------------------
def style_affine(w_lat: torch.Tensor, sd, prefix: str) -> torch.Tensor:
    """stylegan2/modules.py:936 via :880-894: dense (coef 1/sqrt(latent)) + bias (coef 1, init 1)."""
    a = sd[prefix + ".dense.layer.weight"]
    return w_lat.matmul((a * _coef(a.shape)).t()) + sd[prefix + ".dense.bias"]


def modulated_conv(x: torch.Tensor, w_lat: torch.Tensor, sd, prefix: str,
                   demodulate: bool, up: bool) -> torch.Tensor:
    """stylegan2/modules.py:920-967 (forward_mod) followed by :985-994
    (_process) or, for ``up``, :1089-1139 (conv_transpose2d stride 2, then the
    FIR with pad 1 set up at :1049-1072)."""
    weight = sd[prefix + ".weight"]
    B, I = x.shape[0], x.shape[1]
    O, k = weight.shape[0], weight.shape[-1]
    s = style_affine(w_lat, sd, prefix)                              # [B, I]
    w = (weight * _coef(weight.shape))[None] * s[:, None, :, None, None]   # [B,O,I,k,k]
    if demodulate:
        d = torch.rsqrt((w.reshape(B, O, -1) ** 2).sum(-1) + EPS)   # modules.py:945-954
        w = w * d[:, :, None, None, None]
    xg = x.reshape(1, B * I, *x.shape[2:])
    if up:
        wt = w.transpose(1, 2).reshape(B * I, O, k, k)               # modules.py:141-166
        y = F.conv_transpose2d(xg, wt, stride=2, groups=B)           # (2H+1)^2
        y = fir(y, fir_kernel(1.0, 2), 1, 1)                         # -> (2H)^2
    else:
        y = F.conv2d(xg, w.reshape(B * O, I, k, k), padding=k // 2, groups=B)
    return y.reshape(B, O, *y.shape[2:])


def upsample_skip(y: torch.Tensor) -> torch.Tensor:
    """stylegan2/modules.py:580-602: zero-insert via conv_transpose2d(ones,
    stride 2) -> (2H-1)^2, pad [3,1,3,1] (:569-576), FIR (gain 1, up 2)."""
    c = y.shape[1]
    u = F.conv_transpose2d(y, torch.ones(c, 1, 1, 1), stride=2, groups=c)
    return fir(u, fir_kernel(1.0, 2), 3, 1)


# ----------------------------------------------------------------------------
# synthesis network
# ----------------------------------------------------------------------------
def synthesis(w_lat: torch.Tensor, sd, num_blocks: int,
              noise: Optional[Sequence[torch.Tensor]] = None,
              capture: Optional[dict] = None) -> torch.Tensor:
    """stylegan2/models.py:969-1014 with the block body of
    stylegan2/modules.py:1412-1436 and the wrapper order
    conv -> noise (:414-453) -> bias/act (:276-297).

    ``w_lat`` is [B, latent]; clip-glass feeds the same dlatent to all style
    layers (stylegan2/models.py:427-430).  ``noise`` is the list of
    [1,1,H,W] tensors for this minibatch (None => no noise added).
    """
    B = w_lat.shape[0]
    x = sd["G_synthesis.const"][None].expand(B, -1, -1, -1)
    y = None
    ni = 0
    for b in range(num_blocks):
        nl = 1 if b == 0 else 2
        for l in range(nl):
            p = f"G_synthesis.conv_blocks.{b}.conv_block.{l}"
            x = modulated_conv(x, w_lat, sd, p + ".layer.layer", demodulate=True,
                               up=(b > 0 and l == 0))
            if noise is not None:
                x = x + sd[p + ".layer.weight"] * noise[ni]         # modules.py:452
            ni += 1
            x = lrelu_gain(x + sd[p + ".bias"].view(1, -1, 1, 1))
            if capture is not None:
                capture[f"b{b}l{l}"] = x
        if y is not None:
            y = upsample_skip(y)                                    # models.py:1004-1006
        p = f"G_synthesis.to_data_layers.{b}"
        t = modulated_conv(x, w_lat, sd, p + ".layer", demodulate=False, up=False)
        t = t + sd[p + ".bias"].view(1, -1, 1, 1)                   # linear act, gain 1
        y = t if y is None else y + t                               # models.py:1011-1013
        if capture is not None:
            capture[f"rgb{b}"] = y
    return y


def generator(z: torch.Tensor, sd, num_blocks: int, mapping_layers: int = 8,
              noise=None, capture=None) -> torch.Tensor:
    """stylegan2/models.py:326-482 as clip-glass uses it: mapping -> same
    dlatent for every layer -> truncation is the identity because
    ``layer_psi`` is None after ``load()`` (models.py:255-256, 322-324) ->
    synthesis."""
    return synthesis(mapping(z, sd, mapping_layers), sd, num_blocks, noise, capture)


# ----------------------------------------------------------------------------
# discriminator
# ----------------------------------------------------------------------------
def minibatch_std(x: torch.Tensor, group_size: int = 4) -> torch.Tensor:
    """stylegan2/modules.py:701-747.

    QUIRK (replicated, not fixed): for fp32 input, ``y = input.view(...)`` at
    :726 aliases the input and ``y.float()`` at :728 is a no-op, so the
    in-place ``y -= y.mean(dim=0)`` at :730 ALSO subtracts the group mean from
    the features that are concatenated at :746 and fed to the final conv.
    The reference's D therefore sees group-centred features; so does this
    restatement (written out-of-place)."""
    B = x.shape[0]
    g = group_size or B
    y = x.reshape(g, -1, *x.shape[1:]).float()
    y = y - y.mean(dim=0, keepdim=True)
    centred = y.reshape(x.shape).to(x)
    y = torch.sqrt((y ** 2).mean(dim=0) + EPS)
    y = y.reshape(y.shape[0], -1).mean(dim=-1)
    y = y.reshape(-1, 1, 1, 1).repeat(g, 1, 1, 1).expand(B, 1, *x.shape[2:])
    return torch.cat([centred, y.to(x)], dim=1)


def discriminator(img: torch.Tensor, sd, num_blocks: int, group_size: int = 4,
                  capture: Optional[dict] = None) -> torch.Tensor:
    """stylegan2/models.py:1193-1230 with DiscriminatorConvBlock
    (stylegan2/modules.py:1587-1601) and ConvDownLayer (:1238-1254; FIR pad
    from :1204-1209)."""
    w = sd["from_data_layers.0.layer.weight"]
    x = F.conv2d(img, w * _coef(w.shape))
    x = lrelu_gain(x + sd["from_data_layers.0.bias"].view(1, -1, 1, 1))
    blur = fir_kernel(1.0, 1)
    for b in range(num_blocks - 1):
        p = f"conv_blocks.{b}"
        w0 = sd[p + ".conv_block.0.layer.weight"]
        a = F.conv2d(x, w0 * _coef(w0.shape), padding=1)
        a = lrelu_gain(a + sd[p + ".conv_block.0.bias"].view(1, -1, 1, 1))
        w1 = sd[p + ".conv_block.1.layer.weight"]
        a = F.conv2d(fir(a, blur, 2, 2), w1 * _coef(w1.shape), stride=2)
        a = lrelu_gain(a + sd[p + ".conv_block.1.bias"].view(1, -1, 1, 1))
        wp = sd[p + ".projection.weight"]
        r = F.conv2d(fir(x, blur, 1, 1), wp * _coef(wp.shape), stride=2)
        x = (a + r) * (1.0 / SQRT2)
        if capture is not None:
            capture[f"d{b}"] = x
    if group_size:
        x = minibatch_std(x, group_size)
    p = f"conv_blocks.{num_blocks - 1}.1.conv_block.0"
    w = sd[p + ".layer.weight"]
    x = F.conv2d(x, w * _coef(w.shape), padding=1)
    x = lrelu_gain(x + sd[p + ".bias"].view(1, -1, 1, 1))
    x = x.reshape(x.shape[0], -1)
    w = sd["dense.0.layer.weight"]
    x = lrelu_gain(x.matmul((w * _coef(w.shape)).t()) + sd["dense.0.bias"])
    w = sd["dense.1.layer.weight"]
    return x.matmul((w * _coef(w.shape)).t()) + sd["dense.1.bias"]
