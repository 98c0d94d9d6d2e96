"""TextEngine: Python handle on the img2txt C-ABI engine (GPT-2 greedy decode + CLIP text tower; one GPU).

Mirrors what the reference's ``GPT2`` model wrapper (models.py:14-62) and the img2txt branch of
``Generator.clip_similarity`` (generator.py:53-59) compute, at the token level; the string work around it (BPE
decode, clip.tokenize) stays on the host (clip_glass_b200/tokenizers.py) as it does in the reference.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from ._lib import GlassArgError, GlassError, load_library
from .text_weights import ClipTextSpec, GPT2Spec, text_as_built

LO_SCALE = 2048.0
TEXT_FLAG_NO_SPLIT_K = 1        # include/clipglass_b200.h: GLASS_TEXT_FLAG_NO_SPLIT_K


class GlassTextConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("gpt2_vocab", "gpt2_positions", "gpt2_embd", "gpt2_layers", "gpt2_heads")] + \
               [("gpt2_eps", ctypes.c_float)] + \
               [(n, ctypes.c_int32) for n in ("dim_z", "n_init", "max_tokens_len", "text_width", "text_heads", "text_layers",
                                              "text_context", "text_vocab", "text_embed_dim", "max_population", "device",
                                              "flags")]


def split_fp16(w: np.ndarray):
    """x -> (hi, lo) with hi = fp16(x), lo = fp16((x - hi) * 2^11): the split-fp16 operand format of the GPT-2 GEMMs
    (text_engine.cu).  22 significant bits; lo stays in fp16's normal range for |x| > 2^-14."""
    w = np.ascontiguousarray(w, dtype=np.float32)
    hi = w.astype(np.float16)
    lo = ((w - hi.astype(np.float32)) * np.float32(LO_SCALE)).astype(np.float16)
    return hi, lo


def pack_gpt2(sd: Dict[str, torch.Tensor], spec: GPT2Spec, init_tokens: Sequence[int]) -> Dict[str, np.ndarray]:
    """Reference key layout (gpt2/model.py; Conv1D weights are [in, out], x @ W) -> engine tensors.  GEMM weights are
    stored transposed ([out, in], K-major) and split into hi / lo fp16 parts; the tied LM head is the (zero-padded)
    embedding matrix."""
    f = lambda k: sd[k].detach().float().cpu().numpy()
    E = spec.n_embd
    npad = (spec.vocab + 63) // 64 * 64
    out = {"g2.wte": f("transformer.wte.weight"), "g2.wpe": f("transformer.wpe.weight"),
           "g2.init": np.asarray(list(init_tokens), dtype=np.int32),
           "g2.lnf.w": f("transformer.ln_f.weight"), "g2.lnf.b": f("transformer.ln_f.bias")}
    head = np.zeros((npad, E), dtype=np.float32)
    head[:spec.vocab] = out["g2.wte"]
    out["g2.wte.hi"], out["g2.wte.lo"] = split_fp16(head)
    for l in range(spec.n_layer):
        p, q = f"transformer.h.{l}.", f"g2.l{l}."
        for a, b in (("ln_1", "ln1"), ("ln_2", "ln2")):
            out[q + b + ".w"], out[q + b + ".b"] = f(p + a + ".weight"), f(p + a + ".bias")
        for a, b in (("attn.c_attn", "attn"), ("attn.c_proj", "proj"), ("mlp.c_fc", "fc"), ("mlp.c_proj", "proj2")):
            out[q + b + ".w.hi"], out[q + b + ".w.lo"] = split_fp16(f(p + a + ".weight").T)
            out[q + b + ".b"] = f(p + a + ".bias")
    return out


def pack_clip_text(sd: Dict[str, torch.Tensor], spec: ClipTextSpec) -> Dict[str, np.ndarray]:
    """CLIP text half (clip/model.py:277-290 key names) "as built": fp16 Linear / MHA weights, fp32 LayerNorm; the
    embeddings are cast to fp16 where they are used (clip/model.py:308-310), so they are stored fp16."""
    b = text_as_built(sd)
    h = lambda k: np.ascontiguousarray(b[k].detach().half().cpu().numpy())
    f = lambda k: np.ascontiguousarray(b[k].detach().float().cpu().numpy())
    out = {"t.tok": h("token_embedding.weight"), "t.pos": h("positional_embedding"),
           "t.lnf.w": f("ln_final.weight"), "t.lnf.b": f("ln_final.bias"), "t.proj": f("text_projection")}
    for l in range(spec.layers):
        p, q = f"transformer.resblocks.{l}.", f"t.l{l}."
        out[q + "ln1.w"], out[q + "ln1.b"] = f(p + "ln_1.weight"), f(p + "ln_1.bias")
        out[q + "ln2.w"], out[q + "ln2.b"] = f(p + "ln_2.weight"), f(p + "ln_2.bias")
        out[q + "qkv.w"], out[q + "qkv.b"] = h(p + "attn.in_proj_weight"), f(p + "attn.in_proj_bias")
        out[q + "out.w"], out[q + "out.b"] = h(p + "attn.out_proj.weight"), f(p + "attn.out_proj.bias")
        out[q + "fc.w"], out[q + "fc.b"] = h(p + "mlp.c_fc.weight"), f(p + "mlp.c_fc.bias")
        out[q + "proj.w"], out[q + "proj.b"] = h(p + "mlp.c_proj.weight"), f(p + "mlp.c_proj.bias")
    return out


_bound = False


def _bind(lib):
    global _bound
    if _bound:
        return
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    lib.glass_text_create.argtypes = [ctypes.POINTER(GlassTextConfig), ctypes.POINTER(vp)]
    lib.glass_text_set_tensor.argtypes = [vp, ctypes.c_char_p, vp, ctypes.c_size_t]
    lib.glass_text_finalize.argtypes = [vp]
    lib.glass_text_set_image_features.argtypes = [vp, vp, i32]
    lib.glass_text_generate.argtypes = [vp, vp, i32, vp, vp]
    lib.glass_text_similarity.argtypes = [vp, vp, i32, vp, vp, vp]
    lib.glass_text_launch_count.argtypes = [vp]
    lib.glass_text_launch_count.restype = i64
    lib.glass_text_last_error.restype = ctypes.c_char_p
    lib.glass_text_destroy.argtypes = [vp]
    lib.glass_text_set_timing.argtypes = [vp, i32]
    lib.glass_text_gemm_time.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(i32),
                                         ctypes.POINTER(ctypes.c_double)]
    _bound = True


def _check(lib, rc: int) -> None:
    if rc < 0:
        msg = lib.glass_text_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise GlassArgError(msg)
        raise GlassError(f"clipglass_b200 text engine error {rc}: {msg}")


class TextEngine:
    def __init__(self, gpt2: Optional[GPT2Spec], gpt2_sd, text: Optional[ClipTextSpec], text_sd, init_tokens=(),
                 dim_z: int = 20, max_tokens_len: int = 30, max_population: int = 64, device: int = 0, flags: int = 0):
        self.lib = load_library()
        _bind(self.lib)
        self.gpt2, self.text = gpt2, text
        self.dim_z, self.n_init, self.max_tokens_len = dim_z, len(init_tokens), max_tokens_len
        self.max_population, self.device = max_population, device
        cfg = GlassTextConfig()
        if gpt2 is not None:
            cfg.gpt2_vocab, cfg.gpt2_positions, cfg.gpt2_embd = gpt2.vocab, gpt2.n_positions, gpt2.n_embd
            cfg.gpt2_layers, cfg.gpt2_heads, cfg.gpt2_eps = gpt2.n_layer, gpt2.n_head, gpt2.eps
            cfg.dim_z, cfg.n_init, cfg.max_tokens_len = dim_z, len(init_tokens), max_tokens_len
        if text is not None:
            cfg.text_width, cfg.text_heads, cfg.text_layers = text.width, text.heads, text.layers
            cfg.text_context, cfg.text_vocab, cfg.text_embed_dim = text.context, text.vocab, text.embed_dim
        cfg.max_population, cfg.device, cfg.flags = max_population, device, flags
        self._h = ctypes.c_void_p()
        _check(self.lib, self.lib.glass_text_create(ctypes.byref(cfg), ctypes.byref(self._h)))
        packed = {}
        if gpt2 is not None:
            packed.update(pack_gpt2(gpt2_sd, gpt2, init_tokens))
        if text is not None:
            packed.update(pack_clip_text(text_sd, text))
        for name, arr in packed.items():
            arr = np.ascontiguousarray(arr)
            _check(self.lib, self.lib.glass_text_set_tensor(self._h, name.encode(), arr.ctypes.data, arr.nbytes))
        _check(self.lib, self.lib.glass_text_finalize(self._h))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.glass_text_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_image_features(self, image_features) -> None:
        """generator.py:26-27: the cached ``CLIP.encode_image(target)`` [1,E]."""
        t = np.ascontiguousarray(np.asarray(torch.as_tensor(image_features).float().cpu()).reshape(-1), dtype=np.float32)
        _check(self.lib, self.lib.glass_text_set_image_features(self._h, t.ctypes.data, t.size))

    def generate_tokens(self, z: np.ndarray) -> np.ndarray:
        """models.py:45-60 + gpt2/sample.py:21-37 (sample=False): int64 [P, dim_z] -> int64 [P, dim_z + n_init + 30]."""
        z = np.ascontiguousarray(z, dtype=np.int64)
        assert z.ndim == 2 and z.shape[1] == self.dim_z, z.shape
        out = np.empty((z.shape[0], self.dim_z + self.n_init + self.max_tokens_len), dtype=np.int64)
        _check(self.lib, self.lib.glass_text_generate(self._h, z.ctypes.data, z.shape[0], out.ctypes.data, self._stream()))
        return out

    def text_similarity(self, clip_tokens: np.ndarray, return_features: bool = False):
        """generator.py:57-59: int64 [P, context] (clip.tokenize output) -> cosine vs the cached image features."""
        t = np.ascontiguousarray(clip_tokens, dtype=np.int64)
        assert t.ndim == 2 and t.shape[1] == self.text.context, t.shape
        sim = np.empty(t.shape[0], dtype=np.float32)
        feats = np.empty((t.shape[0], self.text.embed_dim), dtype=np.float32) if return_features else None
        _check(self.lib, self.lib.glass_text_similarity(self._h, t.ctypes.data, t.shape[0], sim.ctypes.data,
                                                        feats.ctypes.data if feats is not None else None, self._stream()))
        return (sim, feats) if return_features else sim

    @property
    def launch_count(self) -> int:
        return int(self.lib.glass_text_launch_count(self._h))

    def set_timing(self, enable: bool = True) -> None:
        _check(self.lib, self.lib.glass_text_set_timing(self._h, int(enable)))

    def gemm_time(self):
        """(ms, launches, algorithmic bytes) of the tensor-core GEMM launches since the last call (timing enabled)."""
        ms, n, b = ctypes.c_float(), ctypes.c_int32(), ctypes.c_double()
        _check(self.lib, self.lib.glass_text_gemm_time(self._h, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(b)))
        return float(ms.value), int(n.value), float(b.value)
