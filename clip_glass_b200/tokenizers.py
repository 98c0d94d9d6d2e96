"""Host-side string work of the img2txt path: byte-level BPE for GPT-2 (gpt2/encoder.py, used by models.py:24,30,
32-42) and CLIP's lower-cased BPE (clip/simple_tokenizer.py + clip/clip.py:125-139 ``tokenize``).

The reference keeps this on the CPU too (models.py:32-42 decodes ``.cpu().numpy().tolist()`` token lists;
generator.py:53-57 re-tokenises Python strings), so it is host code here as well, written from the published
algorithm: greedy lowest-rank pair merging over a byte->unicode alphabet.  Vocabulary files are NOT part of this repo:
they are read from the paths the reference's config names (``config.encoder`` / ``config.vocab`` for GPT-2,
``bpe_simple_vocab_16e6.txt.gz`` for CLIP).  tests/test_host_cpu.py checks the GPT-2 codec against the reference's
``Encoder`` on the vocabulary in /root/reference; CLIP's vocabulary file is not in the reference tree (and ftfy is
not installed), so the CLIP tokenizer is unpinned — it is exercised on a synthetic merge table only.
"""
from __future__ import annotations

import gzip
import html
import json
from functools import lru_cache
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np
import regex as re


@lru_cache()
def byte_alphabet() -> Dict[int, str]:
    """The reversible byte -> printable unicode map of GPT-2's BPE: printable latin-1 bytes map to themselves, the
    remaining 68 bytes to code points 256, 257, ... in byte order."""
    keep = list(range(33, 127)) + list(range(161, 173)) + list(range(174, 256))
    table, extra = {}, 0
    for b in range(256):
        if b in keep:
            table[b] = chr(b)
        else:
            table[b] = chr(256 + extra)
            extra += 1
    return table


class BytePairCodec:
    """Greedy BPE over symbol tuples: repeatedly merge the adjacent pair with the lowest rank."""

    def __init__(self, ranks: Dict[Tuple[str, str], int], end_of_word: str = ""):
        self.ranks = ranks
        self.eow = end_of_word
        self._cache: Dict[str, Tuple[str, ...]] = {}

    def merge(self, token: str) -> Tuple[str, ...]:
        hit = self._cache.get(token)
        if hit is not None:
            return hit
        syms = list(token)
        if self.eow and syms:
            syms[-1] += self.eow
        while len(syms) > 1:
            best, best_rank = -1, None
            for i in range(len(syms) - 1):
                r = self.ranks.get((syms[i], syms[i + 1]))
                if r is not None and (best_rank is None or r < best_rank):
                    best, best_rank = i, r
            if best_rank is None:
                break
            a, b = syms[best], syms[best + 1]
            out, i = [], 0
            while i < len(syms):                       # merge EVERY occurrence of the chosen pair, left to right
                if i < len(syms) - 1 and syms[i] == a and syms[i + 1] == b:
                    out.append(a + b)
                    i += 2
                else:
                    out.append(syms[i])
                    i += 1
            syms = out
        res = tuple(syms)
        self._cache[token] = res
        return res


class GPT2Tokenizer:
    """gpt2/encoder.py:41-115 (``get_encoder(config)``): ``encoder.json`` (symbol -> id) + ``vocab.bpe`` (merges)."""

    PATTERN = r"""'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+"""

    def __init__(self, encoder_json: str, vocab_bpe: str):
        with open(encoder_json, "r") as f:
            self.encoder: Dict[str, int] = json.load(f)
        with open(vocab_bpe, "r", encoding="utf-8") as f:
            lines = f.read().split("\n")[1:-1]
        self.decoder = {v: k for k, v in self.encoder.items()}
        self.codec = BytePairCodec({tuple(l.split()): i for i, l in enumerate(lines)})
        self.b2u = byte_alphabet()
        self.u2b = {v: k for k, v in self.b2u.items()}
        self.pat = re.compile(self.PATTERN)
        self.eot = self.encoder["<|endoftext|>"]

    def encode(self, text: str) -> List[int]:
        ids: List[int] = []
        for piece in self.pat.findall(text):
            mapped = "".join(self.b2u[b] for b in piece.encode("utf-8"))
            ids.extend(self.encoder[s] for s in self.codec.merge(mapped))
        return ids

    def decode(self, ids: Iterable[int]) -> str:
        text = "".join(self.decoder[int(i)] for i in ids)
        return bytearray(self.u2b[c] for c in text).decode("utf-8", errors="replace")

    def parse_out(self, seqs: Sequence[Sequence[int]], dim_z: int, max_text_len: int) -> List[str]:
        """models.py:32-42: text = decode(seq[dim_z : first EOT anywhere in seq])[:max_text_len] (the EOT search covers
        the latent genes too — replicated)."""
        texts = []
        for seq in seqs:
            seq = [int(t) for t in seq]
            body = seq[dim_z:seq.index(self.eot)] if self.eot in seq else seq[dim_z:]
            texts.append(self.decode(body)[:max_text_len])
        return texts


def _whitespace_clean(text: str) -> str:
    return re.sub(r"\s+", " ", text).strip()


class ClipTokenizer:
    """clip/simple_tokenizer.py:62-127 + clip/clip.py:125-139.  ``basic_clean`` calls ``ftfy.fix_text`` in the
    reference; ftfy is not installed here, so only its html-unescape half is applied (identity on clean ASCII text)."""

    PATTERN = r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+"""

    def __init__(self, bpe_path: str = None, merges: Sequence[Tuple[str, str]] = None):
        if merges is None:
            lines = gzip.open(bpe_path).read().decode("utf-8").split("\n")
            merges = [tuple(m.split()) for m in lines[1:49152 - 256 - 2 + 1]]
        self.b2u = byte_alphabet()
        self.u2b = {v: k for k, v in self.b2u.items()}
        vocab = list(self.b2u.values())
        vocab = vocab + [v + "</w>" for v in vocab] + ["".join(m) for m in merges] + ["<|startoftext|>", "<|endoftext|>"]
        self.encoder = {s: i for i, s in enumerate(vocab)}
        self.decoder = {i: s for s, i in self.encoder.items()}
        self.codec = BytePairCodec({tuple(m): i for i, m in enumerate(merges)}, end_of_word="</w>")
        self.pat = re.compile(self.PATTERN, re.IGNORECASE)
        self.sot, self.eot = self.encoder["<|startoftext|>"], self.encoder["<|endoftext|>"]

    def encode(self, text: str) -> List[int]:
        text = _whitespace_clean(html.unescape(html.unescape(text)).strip()).lower()
        ids: List[int] = []
        for piece in self.pat.findall(text):
            if piece in ("<|startoftext|>", "<|endoftext|>"):
                ids.append(self.encoder[piece])
                continue
            mapped = "".join(self.b2u[b] for b in piece.encode("utf-8"))
            ids.extend(self.encoder[s] for s in self.codec.merge(mapped))
        return ids

    def decode(self, ids: Iterable[int]) -> str:
        text = "".join(self.decoder[int(i)] for i in ids)
        return bytearray(self.u2b[c] for c in text).decode("utf-8", errors="replace").replace("</w>", " ")

    def tokenize(self, texts, context_length: int = 77) -> np.ndarray:
        """clip/clip.py:125-139: [SOT] + bpe + [EOT], zero padded to ``context_length``; RuntimeError when too long
        (generator.py:55-56 catches it and scores the whole population 0)."""
        if isinstance(texts, str):
            texts = [texts]
        out = np.zeros((len(texts), context_length), dtype=np.int64)
        for i, t in enumerate(texts):
            ids = [self.sot] + self.encode(t) + [self.eot]
            if len(ids) > context_length:
                raise RuntimeError(f"Input {t} is too long for context length {context_length}")
            out[i, :len(ids)] = ids
        return out
