"""ORACLE (test infrastructure only — never on the product path).

CPU restatement of the img2txt fitness path (BASELINE config 5):

  models.py:45-62      GPT2.generate: cat(latent tokens, init tokens) -> sample_sequence -> parse_out
  gpt2/sample.py:21-37 30-step decode with KV cache; stochastic=False (config.py:18) => top-1 of softmax(top-40
                       filtered logits / 0.7), i.e. the arg-max of the raw logits
  gpt2/model.py:45-175 12-layer GPT-2 (TF-style LayerNorm, tanh-GELU, scaled masked attention with `past`)
  generator.py:53-59   clip.tokenize(texts) -> CLIP.encode_text -> cosine vs the cached image features
  clip/model.py:292-320 causal mask, EOT-argmax gather, ln_final, text_projection

Functional torch code over plain state dicts (reference key names), fp32 for GPT-2 (as the reference runs it) and
"fp16 as built" or fp32 arithmetic for the CLIP text tower (clip/model.py:339-360 converts it to fp16).  Pinned by
tests/golden/gpt2_*.npz, which oracle/make_golden_gpt2.py writes from the UNMODIFIED reference modules.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------
# GPT-2 (gpt2/model.py)
# ---------------------------------------------------------------------------
def gpt2_layernorm(x, w, b, eps):
    """gpt2/model.py:16-29: epsilon inside the square root, biased variance."""
    u = x.mean(-1, keepdim=True)
    s = (x - u).pow(2).mean(-1, keepdim=True)
    return w * ((x - u) / torch.sqrt(s + eps)) + b


def gelu_tanh(x):
    """gpt2/model.py:13-14."""
    return 0.5 * x * (1 + torch.tanh(math.sqrt(2 / math.pi) * (x + 0.044715 * torch.pow(x, 3))))


def gpt2_forward(sd: Dict[str, torch.Tensor], spec, tokens: torch.Tensor, cache: Optional[list]):
    """One call of GPT2LMHeadModel.forward (gpt2/model.py:126-175, 196-210) on ``tokens`` [B, T_new] with the KV
    cache of the previous calls (list per layer of (K [B,H,T,64], V [B,H,T,64]) or None).  Returns (logits of the
    LAST position [B, vocab], new cache)."""
    B, Tn = tokens.shape
    H, D = spec.n_head, spec.n_embd // spec.n_head
    past = 0 if cache is None else cache[0][0].shape[2]
    pos = torch.arange(past, past + Tn)
    h = sd["transformer.wte.weight"][tokens] + sd["transformer.wpe.weight"][pos][None]     # :148-156
    new_cache = []
    for l in range(spec.n_layer):
        p = f"transformer.h.{l}."
        a = gpt2_layernorm(h, sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], spec.eps)
        qkv = a @ sd[p + "attn.c_attn.weight"] + sd[p + "attn.c_attn.bias"]                # Conv1D :31-43
        q, k, v = qkv.split(spec.n_embd, dim=2)
        q = q.view(B, Tn, H, D).permute(0, 2, 1, 3)
        k = k.view(B, Tn, H, D).permute(0, 2, 1, 3)
        v = v.view(B, Tn, H, D).permute(0, 2, 1, 3)
        if cache is not None:
            k = torch.cat((cache[l][0], k), dim=2)                                         # :86-89
            v = torch.cat((cache[l][1], v), dim=2)
        new_cache.append((k, v))
        w = (q @ k.transpose(-1, -2)) / math.sqrt(D)                                       # :59-62 (scale=True)
        ns = k.shape[2]
        mask = torch.tril(torch.ones(ns, ns))[ns - Tn:ns, :ns]                             # :63-65
        w = w * mask - 1e10 * (1 - mask)
        w = torch.softmax(w, dim=-1)
        o = (w @ v).permute(0, 2, 1, 3).reshape(B, Tn, spec.n_embd)                        # merge_heads
        h = h + (o @ sd[p + "attn.c_proj.weight"] + sd[p + "attn.c_proj.bias"])
        m = gpt2_layernorm(h, sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], spec.eps)
        m = gelu_tanh(m @ sd[p + "mlp.c_fc.weight"] + sd[p + "mlp.c_fc.bias"])
        h = h + (m @ sd[p + "mlp.c_proj.weight"] + sd[p + "mlp.c_proj.bias"])
    h = gpt2_layernorm(h, sd["transformer.ln_f.weight"], sd["transformer.ln_f.bias"], spec.eps)
    logits = h[:, -1, :] @ sd["transformer.wte.weight"].t()                                # tied head :177-194
    return logits, new_cache


def gpt2_generate_tokens(sd, spec, z: np.ndarray, init_tokens: List[int], length: int, return_logits: bool = False):
    """models.py:45-60 + gpt2/sample.py:21-37 with sample=False: context = cat(z, init_tokens) [P, dim_z + n_init];
    ``length`` greedy steps.  Temperature 0.7, the top-40 filter and the softmax are monotone, so top-1 of the
    probabilities is the arg-max of the logits.  Returns int64 [P, dim_z + n_init + length]."""
    with torch.no_grad():
        ctx = torch.cat((torch.as_tensor(z).long(), torch.tensor(init_tokens).long().repeat(len(z), 1)), dim=1)
        out, prev, cache, first = ctx, ctx, None, None
        for i in range(length):
            logits, cache = gpt2_forward(sd, spec, prev, cache)
            if first is None:
                first = logits
            prev = torch.argmax(logits, dim=-1, keepdim=True)
            out = torch.cat((out, prev), dim=1)
    return (out.numpy(), first.numpy()) if return_logits else out.numpy()


def parse_out_tokens(seqs: np.ndarray, dim_z: int, eot: int) -> List[List[int]]:
    """The token part of models.py:32-42 ``parse_out``: the generated text is seq[dim_z : first EOT] — NB
    ``seq.index(EOT)`` searches the WHOLE sequence, latent genes included, so an EOT gene at position j < dim_z gives
    the empty text (replicated, not fixed)."""
    out = []
    for seq in np.asarray(seqs).tolist():
        out.append(seq[dim_z:seq.index(eot)] if eot in seq else seq[dim_z:])
    return out


# ---------------------------------------------------------------------------
# CLIP text tower (clip/model.py:292-320)
# ---------------------------------------------------------------------------
def clip_encode_text(sd: Dict[str, torch.Tensor], spec, tokens: torch.Tensor, mode: str = "as_built"):
    """``CLIP.encode_text``.  mode "as_built": fp16 parameters and activations with fp32 LayerNorm, the op boundaries
    of the reference (torch CPU half ops); "fp32": the same fp16-rounded parameters, fp32 arithmetic (the "before the
    final cast" comparison value, like clip_oracle's fp32 mode for the image tower)."""
    dt = torch.float16 if mode == "as_built" else torch.float32
    c = lambda t: t.to(dt)
    W, H, L = spec.width, spec.heads, spec.context
    with torch.no_grad():
        x = c(sd["token_embedding.weight"][tokens]) + c(sd["positional_embedding"])       # :308-310
        mask = torch.full((L, L), float("-inf")).triu_(1).to(dt)                           # :292-298
        for l in range(spec.layers):
            p = f"transformer.resblocks.{l}."
            h = F.layer_norm(x.float(), (W,), sd[p + "ln_1.weight"].float(), sd[p + "ln_1.bias"].float(), 1e-5).to(dt)
            qkv = F.linear(h, c(sd[p + "attn.in_proj_weight"]), c(sd[p + "attn.in_proj_bias"]))
            q, k, v = qkv.split(W, dim=-1)
            B = x.shape[0]
            q = q.view(B, L, H, W // H).transpose(1, 2) * ((W // H) ** -0.5)
            k = k.view(B, L, H, W // H).transpose(1, 2)
            v = v.view(B, L, H, W // H).transpose(1, 2)
            a = torch.softmax((q @ k.transpose(-1, -2)) + mask, dim=-1)
            o = (a @ v).transpose(1, 2).reshape(B, L, W)
            x = x + F.linear(o, c(sd[p + "attn.out_proj.weight"]), c(sd[p + "attn.out_proj.bias"]))
            h = F.layer_norm(x.float(), (W,), sd[p + "ln_2.weight"].float(), sd[p + "ln_2.bias"].float(), 1e-5).to(dt)
            h = F.linear(h, c(sd[p + "mlp.c_fc.weight"]), c(sd[p + "mlp.c_fc.bias"]))
            h = h * torch.sigmoid(1.702 * h)
            x = x + F.linear(h, c(sd[p + "mlp.c_proj.weight"]), c(sd[p + "mlp.c_proj.bias"]))
        x = F.layer_norm(x.float(), (W,), sd["ln_final.weight"].float(), sd["ln_final.bias"].float(), 1e-5).to(dt)
        eot = tokens.argmax(dim=-1)                                                        # :318
        return x[torch.arange(x.shape[0]), eot] @ c(sd["text_projection"])


def text_similarity(text_features: torch.Tensor, image_features: torch.Tensor) -> torch.Tensor:
    """generator.py:59: torch.cosine_similarity(text_features, image_features) with image_features [1, E]."""
    return torch.cosine_similarity(text_features.float(), image_features.float())
