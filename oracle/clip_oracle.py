"""ORACLE (test infrastructure only — never on the product path).

CPU restatement of CLIP ViT-B/32 ``encode_image`` as CLIP-GLaSS calls it
(generator.py:49 -> clip/model.py:304-305 -> VisualTransformer.forward
:218-235).  Operates on the visual-tower state dict (``visual.`` prefix
stripped).  Two numeric modes:

  * ``as_built``  — dtypes exactly as the reference builds them
    (clip/model.py:339-360,397): fp16 weights/activations, fp32 LayerNorm
    (clip/model.py:152-158).  This is what the reference returns.
  * ``fp32``      — the same fp16-rounded weights, all arithmetic in fp32.
    Used for tight comparisons "before the final cast" (SURVEY.md §7 hard
    part 2).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F


def _ln(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """clip/model.py:152-158: LayerNorm in fp32, cast back."""
    return F.layer_norm(x.float(), (x.shape[-1],), w.float(), b.float(), 1e-5).to(x.dtype)


def _resblock(x: torch.Tensor, sd: Dict[str, torch.Tensor], p: str, heads: int) -> torch.Tensor:
    """clip/model.py:166-187: x + MHA(ln_1 x); x + c_proj(QuickGELU(c_fc(ln_2 x))).
    x is [L, N, D] (clip/model.py:226)."""
    dt = x.dtype
    h = _ln(x, sd[p + ".ln_1.weight"], sd[p + ".ln_1.bias"])
    a, _ = F.multi_head_attention_forward(      # what nn.MultiheadAttention.forward calls
        h, h, h, h.shape[-1], heads,
        sd[p + ".attn.in_proj_weight"].to(dt), sd[p + ".attn.in_proj_bias"].to(dt),
        None, None, False, 0.0,
        sd[p + ".attn.out_proj.weight"].to(dt), sd[p + ".attn.out_proj.bias"].to(dt),
        training=False, need_weights=False, attn_mask=None)
    x = x + a
    h = _ln(x, sd[p + ".ln_2.weight"], sd[p + ".ln_2.bias"])
    h = F.linear(h, sd[p + ".mlp.c_fc.weight"].to(dt), sd[p + ".mlp.c_fc.bias"].to(dt))
    h = h * torch.sigmoid(1.702 * h)                                   # clip/model.py:161-163
    h = F.linear(h, sd[p + ".mlp.c_proj.weight"].to(dt), sd[p + ".mlp.c_proj.bias"].to(dt))
    return x + h


def encode_image(image: torch.Tensor, sd: Dict[str, torch.Tensor], layers: int, heads: int,
                 mode: str = "as_built", capture=None) -> torch.Tensor:
    """image [N,3,R,R] fp32 -> features [N, embed].  ``sd`` holds the as-built
    (fp16 / fp32) tensors from ``weights.clip_as_built``."""
    assert mode in ("as_built", "fp32")
    dt = sd["conv1.weight"].dtype if mode == "as_built" else torch.float32
    x = image.to(dt)                                                   # clip/model.py:305
    patch = sd["conv1.weight"].shape[-1]
    x = F.conv2d(x, sd["conv1.weight"].to(dt), stride=patch)           # :219
    x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)         # :220-221
    cls = sd["class_embedding"].to(dt) + torch.zeros(x.shape[0], 1, x.shape[-1], dtype=dt)
    x = torch.cat([cls, x], dim=1)                                     # :222
    x = x + sd["positional_embedding"].to(dt)                          # :223
    x = _ln(x, sd["ln_pre.weight"], sd["ln_pre.bias"])                 # :224
    if capture is not None:
        capture["ln_pre"] = x
    x = x.permute(1, 0, 2)                                             # :226
    for l in range(layers):
        x = _resblock(x, sd, f"transformer.resblocks.{l}", heads)
        if capture is not None:
            capture[f"block{l}"] = x.permute(1, 0, 2)
    x = x.permute(1, 0, 2)                                             # :228
    x = _ln(x[:, 0, :], sd["ln_post.weight"], sd["ln_post.bias"])      # :230
    return x @ sd["proj"].to(dt)                                       # :232-233
