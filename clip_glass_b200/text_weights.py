"""Specs and seeded synthetic weights of the img2txt path (GPT-2 + CLIP text tower) in the reference's state_dict
layouts: ``gpt2.model.GPT2LMHeadModel`` (gpt2/model.py:126-210; key names ``transformer.wte.weight``,
``transformer.h.{l}.attn.c_attn.weight`` [n_embd, 3 n_embd] (Conv1D: x @ W), ...) and the text half of
``clip.model.CLIP`` (clip/model.py:277-290: ``token_embedding.weight``, ``positional_embedding``,
``transformer.resblocks.{l}.*``, ``ln_final.*``, ``text_projection``).  There are no checkpoints offline; the same
dicts load into the reference modules (oracle/make_golden_gpt2.py) and into the engine.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import torch


@dataclass(frozen=True)
class GPT2Spec:                     # gpt2/config.py:6-24
    vocab: int = 50257
    n_positions: int = 1024
    n_embd: int = 768
    n_layer: int = 12
    n_head: int = 12
    eps: float = 1e-5


@dataclass(frozen=True)
class ClipTextSpec:                 # clip/model.py:363-392 for ViT-B/32: width 512, 8 heads, 12 layers
    width: int = 512
    heads: int = 8
    layers: int = 12
    context: int = 77
    vocab: int = 49408
    embed_dim: int = 512


GPT2_SMALL = GPT2Spec()
CLIP_TEXT_B32 = ClipTextSpec()
# reduced shapes for fast tests: same code paths (head dim 64 in both towers)
TINY_GPT2 = GPT2Spec(vocab=4096, n_positions=64, n_embd=128, n_layer=2, n_head=2)
TINY_CLIP_TEXT = ClipTextSpec(width=128, heads=2, layers=2, vocab=8192)


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def make_gpt2_weights(spec: GPT2Spec = GPT2_SMALL, seed: int = 0) -> Dict[str, torch.Tensor]:
    """fp32 state dict; init scale of the reference (normal 0.02, gpt2/model.py:34-37) with non-trivial LayerNorm
    gains / biases so that those code paths are exercised."""
    g = _gen(seed)
    n = lambda *s, std=0.02: torch.randn(*s, generator=g) * std
    E = spec.n_embd
    sd = {"transformer.wte.weight": n(spec.vocab, E), "transformer.wpe.weight": n(spec.n_positions, E, std=0.01)}
    for l in range(spec.n_layer):
        p = f"transformer.h.{l}."
        sd[p + "ln_1.weight"] = 1.0 + n(E, std=0.1)
        sd[p + "ln_1.bias"] = n(E, std=0.1)
        sd[p + "attn.c_attn.weight"] = n(E, 3 * E, std=0.05)
        sd[p + "attn.c_attn.bias"] = n(3 * E)
        sd[p + "attn.c_proj.weight"] = n(E, E, std=0.05)
        sd[p + "attn.c_proj.bias"] = n(E)
        sd[p + "ln_2.weight"] = 1.0 + n(E, std=0.1)
        sd[p + "ln_2.bias"] = n(E, std=0.1)
        sd[p + "mlp.c_fc.weight"] = n(E, 4 * E, std=0.05)
        sd[p + "mlp.c_fc.bias"] = n(4 * E)
        sd[p + "mlp.c_proj.weight"] = n(4 * E, E, std=0.05)
        sd[p + "mlp.c_proj.bias"] = n(E)
    sd["transformer.ln_f.weight"] = 1.0 + n(E, std=0.1)
    sd["transformer.ln_f.bias"] = n(E, std=0.1)
    return sd


def normalise_gpt2_keys(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """``gpt2-pytorch_model.bin`` stores TF-style names without the ``transformer.`` prefix and with ``.g/.b/.w``
    suffixes; gpt2/utils.py:10-52 ``load_weight`` renames them.  Same mapping here."""
    out = {}
    for k, v in sd.items():
        if k.endswith(".g") or k.endswith(".w"):
            k = k[:-2] + ".weight"
        elif k.endswith(".b"):
            k = k[:-2] + ".bias"
        if not k.startswith("transformer.") and not k.startswith("lm_head."):
            k = "transformer." + k
        out[k] = v
    return out


def make_clip_text_weights(spec: ClipTextSpec = CLIP_TEXT_B32, seed: int = 0) -> Dict[str, torch.Tensor]:
    """fp32 master weights of CLIP's text half (init as clip/model.py initialises a fresh model, plus non-trivial
    LayerNorm parameters); ``text_as_built`` applies the fp16 conversion."""
    g = _gen(seed)
    n = lambda *s, std=1.0: torch.randn(*s, generator=g) * std
    W, Ly = spec.width, spec.layers
    sd = {"token_embedding.weight": n(spec.vocab, W, std=0.02), "positional_embedding": n(spec.context, W, std=0.01)}
    attn_std, proj_std, fc_std = W ** -0.5, (W ** -0.5) * ((2 * Ly) ** -0.5), (2 * W) ** -0.5
    for l in range(Ly):
        p = f"transformer.resblocks.{l}."
        sd[p + "attn.in_proj_weight"] = n(3 * W, W, std=attn_std)
        sd[p + "attn.in_proj_bias"] = n(3 * W, std=0.02)
        sd[p + "attn.out_proj.weight"] = n(W, W, std=proj_std)
        sd[p + "attn.out_proj.bias"] = n(W, std=0.02)
        sd[p + "ln_1.weight"] = 1.0 + n(W, std=0.1)
        sd[p + "ln_1.bias"] = n(W, std=0.1)
        sd[p + "mlp.c_fc.weight"] = n(4 * W, W, std=fc_std)
        sd[p + "mlp.c_fc.bias"] = n(4 * W, std=0.02)
        sd[p + "mlp.c_proj.weight"] = n(W, 4 * W, std=proj_std)
        sd[p + "mlp.c_proj.bias"] = n(W, std=0.02)
        sd[p + "ln_2.weight"] = 1.0 + n(W, std=0.1)
        sd[p + "ln_2.bias"] = n(W, std=0.1)
    sd["ln_final.weight"] = 1.0 + n(W, std=0.1)
    sd["ln_final.bias"] = n(W, std=0.1)
    sd["text_projection"] = n(W, spec.embed_dim, std=W ** -0.5)
    return sd


def text_as_built(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """clip/model.py:339-360 ``convert_weights``: Linear / MultiheadAttention parameters and ``text_projection``
    become fp16; LayerNorm, the token embedding (nn.Embedding is not converted) and the positional embedding stay
    fp32 (both are cast with ``.type(self.dtype)`` where they are used, clip/model.py:308-310)."""
    keep32 = ("ln_", "token_embedding", "positional_embedding")
    return {k: (v.float() if any(t in k for t in keep32) else v.half()) for k, v in sd.items()}


def make_token_latents(pop: int, dim_z: int, vocab: int, seed: int):
    """Population as pymoo hands it to ``_evaluate`` for config GPT2: integer genes in [0, vocab) (config.py:22-27,
    int_random sampling)."""
    import numpy as np
    return np.random.default_rng(seed).integers(0, vocab, size=(pop, dim_z)).astype(np.int64)
