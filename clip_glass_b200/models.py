"""Mirror of the reference's ``GPT2`` model wrapper (models.py:14-62) on top of the B200 text engine.

``generate(z, minibatch)`` keeps the reference's contract: ``z`` int64 [P, dim_z] token latents -> list of P strings
(``parse_out``: decode ``seq[dim_z : first EOT]``, truncated to ``max_text_len`` characters).  As in the reference the
whole population is decoded at once and ``minibatch`` is ignored (models.py:46 "TODO: implement minibatch").  The
decode itself (30 greedy steps with a KV cache over P x 53 tokens) runs in ``glass_text_generate``; only the token ->
string step is host code, because it is host code in the reference too.
"""
from __future__ import annotations

import os
from typing import List, Optional

import numpy as np
import torch

from . import text_weights as TW
from ._lib import GlassError
from .text_engine import TextEngine
from .tokenizers import GPT2Tokenizer

INIT_TEXT_TOKENS = {"the picture of": [1169, 4286, 286]}      # gpt2 BPE of config.init_text (config.py:15)


def standin_clip_tokens(gen_tokens: List[List[int]], text_spec: TW.ClipTextSpec) -> np.ndarray:
    """Token-level stand-in for ``clip.tokenize(parse_out(...))`` when no vocabulary files exist (this repo ships
    none; the GPU box has none): SOT, the generated GPT-2 tokens (up to context - 2) mapped into CLIP's id range below
    SOT, EOT, zero padding — the row shape clip/clip.py:125-139 produces.  Used by bench.py's gpt2 workload and the
    GPU tests; a run with real vocabularies goes through tokenizers.ClipTokenizer instead."""
    sot, eot = text_spec.vocab - 2, text_spec.vocab - 1
    out = np.zeros((len(gen_tokens), text_spec.context), dtype=np.int64)
    for i, toks in enumerate(gen_tokens):
        body = [(int(t) * 7 + 13) % (text_spec.vocab - 258) + 256 for t in toks[:text_spec.context - 2]]
        row = [sot] + body + [eot]
        out[i, :len(row)] = row
    return out


class GPT2:
    def __init__(self, config, engine: Optional[TextEngine] = None, text_spec=None, text_sd=None):
        self.config = config
        self.spec = getattr(config, "gpt2_spec", TW.GPT2_SMALL)
        seed = getattr(config, "synthetic_seed", None)
        have_real = isinstance(getattr(config, "weights", None), str) and os.path.exists(config.weights)
        if seed is None and not have_real:
            raise GlassError("Weights not found!\nRun: ./download-weights.sh GPT2   (models.py:18-20); "
                             "or give config.synthetic_seed")
        if have_real and seed is None:
            sd = TW.normalise_gpt2_keys(torch.load(config.weights, map_location="cpu"))       # models.py:22,26
        else:
            sd = TW.make_gpt2_weights(self.spec, seed)
        enc_path, voc_path = getattr(config, "encoder", ""), getattr(config, "vocab", "")
        self.enc = GPT2Tokenizer(enc_path, voc_path) if os.path.exists(enc_path) and os.path.exists(voc_path) else None
        if self.enc is not None:
            init = self.enc.encode(config.init_text)                                          # models.py:30
        elif config.init_text in INIT_TEXT_TOKENS:
            init = INIT_TEXT_TOKENS[config.init_text]
        else:
            raise GlassError(f"no GPT-2 vocabulary at {enc_path!r}: cannot encode init_text {config.init_text!r}")
        self.init_tokens = [t % self.spec.vocab for t in init]
        dev = torch.device(config.device)
        self.engine = engine or TextEngine(
            self.spec, sd, text_spec, text_sd, init_tokens=self.init_tokens, dim_z=config.dim_z,
            max_tokens_len=config.max_tokens_len,
            max_population=int(getattr(config, "max_population", max(config.pop_size, config.batch_size))),
            device=dev.index if dev.index is not None else torch.cuda.current_device())
        self.eot = self.enc.eot if self.enc is not None else self.spec.vocab - 1
        self.last_tokens = None

    def has_discriminator(self):
        return False

    def parse_out_tokens(self, out) -> List[List[int]]:
        seqs = np.asarray(out).tolist()
        return [s[self.config.dim_z:s.index(self.eot)] if self.eot in s else s[self.config.dim_z:] for s in seqs]

    def parse_out(self, out) -> List[str]:                                                    # models.py:32-42
        if self.enc is None:
            raise GlassError("GPT-2 vocabulary files are needed to turn tokens into text (config.encoder / config.vocab)")
        return self.enc.parse_out(np.asarray(out).tolist(), self.config.dim_z, self.config.max_text_len)

    def generate_tokens(self, z) -> np.ndarray:
        z = z.detach().cpu().numpy() if isinstance(z, torch.Tensor) else np.asarray(z)
        self.last_tokens = self.engine.generate_tokens(z.astype(np.int64))
        return self.last_tokens

    def generate(self, z, minibatch=None):                                                    # models.py:45-62
        return self.parse_out(self.generate_tokens(z))
