"""Mirror of the reference's ``Generator`` façade (generator.py:11-71) on top of
the B200 engine.  Same method names, argument meaning and return shapes:

  generate(ls, minibatch)        -> Tensor[P,3,R,R] in [0,1]   (generator.py:29-34)
  discriminate(images, minibatch)-> Tensor[P,1]                (generator.py:36-38)
  has_discriminator()            -> bool                       (generator.py:40-41)
  clip_similarity(images)        -> Tensor[P]                  (generator.py:43-51)
  save(input, path)                                            (generator.py:63-68)

Differences that are deliberate and documented in DESIGN.md:
  * ``clip_similarity`` returns fp32 (the reference returns the fp16-rounded
    value); cast with ``.half()`` for bit-compat.
  * the cached text embedding is supplied as ``config.text_features`` (the
    reference computes it once with CLIP's text tower, generator.py:23-24 —
    outside the per-generation path).
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch

from . import weights as W
from .engine import GlassEngine
from ._lib import GlassArgError, GlassError


def _load_state_dicts(config):
    """G.pth / D.pth are the reference's pickled dicts (stylegan2/models.py:111-132:
    {'name','kwargs','state_dict',...}); ViT-B-32.pt is a TorchScript archive
    whose state_dict has 'visual.*' keys (clip/clip.py:65,77)."""
    gan = getattr(config, "gan_spec", W.FFHQ)
    clip = getattr(config, "clip_spec", W.VIT_B32)
    seed = getattr(config, "synthetic_seed", None)
    wdir = getattr(config, "weights", "")
    have_real = isinstance(wdir, str) and os.path.exists(os.path.join(wdir, "G.pth"))
    if seed is None and not have_real:
        raise GlassError(f"weights not found under {wdir!r} and no synthetic_seed given "
                         "(the reference would print 'Run: ./download-weights.sh' and exit, models.py:91-101)")
    if have_real and seed is None:
        g_sd, g_kw = W.load_reference_checkpoint(os.path.join(wdir, "G.pth"))
        d_sd, d_kw = (W.load_reference_checkpoint(os.path.join(wdir, "D.pth"))
                      if config.use_discriminator else (None, None))
        if not hasattr(config, "gan_spec"):
            gan = W.gan_spec_from_checkpoint(g_sd, g_kw, d_kw)       # church / car / ffhq: built from the pickle
        clip_path = getattr(config, "clip_weights", os.path.expanduser("~/.cache/clip/ViT-B-32.pt"))
        full = torch.jit.load(clip_path, map_location="cpu").state_dict()
        c_sd = {k[len("visual."):]: v for k, v in full.items() if k.startswith("visual.")}
    else:
        g_sd = W.make_generator_weights(gan, seed + 0)
        d_sd = W.make_discriminator_weights(gan, seed + 1) if config.use_discriminator else None
        c_sd = W.make_clip_visual_weights(clip, seed + 2)
    return gan, clip, g_sd, d_sd, c_sd


class Generator:
    def __init__(self, config):
        self.config = config
        self.augmentation = None
        if not str(config.device).startswith("cuda"):
            raise GlassError("the B200 path has no CPU fallback: config.device must be a CUDA device")
        if config.task == "img2txt":
            self._init_img2txt(config)
            return
        if not str(config.device).startswith("cuda"):
            raise GlassError("the B200 path has no CPU fallback: config.device must be a CUDA device")
        dev = torch.device(config.device)
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        gan, clip, g_sd, d_sd, c_sd = _load_state_dicts(config)
        self.gan, self.clip = gan, clip
        max_pop = int(getattr(config, "max_population", max(config.pop_size, config.batch_size)))
        max_pop = (max_pop + config.batch_size - 1) // config.batch_size * config.batch_size
        self.engine = GlassEngine(gan, clip, g_sd, d_sd, c_sd, batch_size=config.batch_size,
                                  max_population=max_pop, device=index,
                                  conv_impl=int(getattr(config, "conv_impl", 0)),
                                  flags=int(getattr(config, "engine_flags", 0)))
        tf = getattr(config, "text_features", None)
        if tf is None:
            raise GlassError("config.text_features ([1,512], CLIP.encode_text of the target) is required")
        self.text_features = torch.as_tensor(tf)
        self.engine.set_text_features(self.text_features)
        self._calls = 0

    # -- img2txt (BASELINE config 5; generator.py:26-27, 53-59, 69-71) -------
    def _init_img2txt(self, config):
        from . import text_weights as TW
        from .models import GPT2
        from .tokenizers import ClipTokenizer
        self.text_spec = getattr(config, "clip_text_spec", TW.CLIP_TEXT_B32)
        seed = getattr(config, "synthetic_seed", None)
        clip_path = getattr(config, "clip_weights", os.path.expanduser("~/.cache/clip/ViT-B-32.pt"))
        if seed is None and os.path.exists(clip_path):
            full = torch.jit.load(clip_path, map_location="cpu").state_dict()
            t_sd = {k: v for k, v in full.items() if not k.startswith("visual.") and k not in
                    ("logit_scale", "input_resolution", "context_length", "vocab_size")}
        elif seed is not None:
            t_sd = TW.make_clip_text_weights(self.text_spec, seed + 1)
        else:
            raise GlassError(f"CLIP weights not found at {clip_path!r} and no synthetic_seed given")
        self.model = GPT2(config, text_spec=self.text_spec, text_sd=t_sd)
        self.engine = self.model.engine
        bpe = getattr(config, "clip_bpe", "")
        self.clip_tokenizer = ClipTokenizer(bpe) if isinstance(bpe, str) and os.path.exists(bpe) else None
        feats = getattr(config, "image_features", None)
        if feats is None:
            raise GlassError("config.image_features ([1,512], CLIP.encode_image of the target picture, "
                             "generator.py:26-27) is required")
        self.image_features = torch.as_tensor(feats)
        self.engine.set_image_features(self.image_features)

    def _text_similarity(self, input):
        """generator.py:53-59.  ``input``: the list of strings ``generate`` returned, or (token-level callers: tests,
        bench.py on a box without vocabulary files) an int64 array of clip tokens [P, context]."""
        if isinstance(input, (np.ndarray, torch.Tensor)):
            tokens = np.asarray(input.cpu() if isinstance(input, torch.Tensor) else input, dtype=np.int64)
        else:
            if self.clip_tokenizer is None:
                raise GlassError("CLIP's BPE vocabulary (config.clip_bpe) is needed to tokenise text")
            try:
                tokens = self.clip_tokenizer.tokenize(list(input), self.text_spec.context)
            except RuntimeError:
                return torch.zeros(len(input))                       # generator.py:55-56
        return torch.from_numpy(self.engine.text_similarity(tokens))

    # -- image output path -------------------------------
    def remember_population(self, x, offset: int = 0):
        """Called by ``GenerationProblem._evaluate`` after a fused evaluation: the engine still holds the images it
        scored (fp32 [P,3,R,R] in its workspace).  run.py:29-51 ``save_callback`` asks for images of candidates of
        the current population — most of them rendered by this very evaluation — through ``generate``; rows that are
        found here are copied out of the engine instead of being rendered a second time."""
        self._last_x32 = np.ascontiguousarray(np.asarray(x, dtype=np.float64)).astype(np.float32)
        self._last_rows = {row.tobytes(): i for i, row in enumerate(self._last_x32)}
        self.reuse_stats = dict(reused=0, rendered=0)

    def forget_population(self):
        """The engine's cached images no longer belong to rows the host has seen (the GPU-resident GA evaluates
        offspring that never reach the host, device_ga.py): ``generate`` renders instead of reusing."""
        self._last_rows = None

    def _cached_rows(self, z_host: np.ndarray):
        rows = getattr(self, "_last_rows", None)
        if not rows or not getattr(self.config, "reuse_evaluated_images", True):
            return [None] * len(z_host)
        return [rows.get(np.ascontiguousarray(r).tobytes()) for r in z_host]

    def _render(self, z, mb: int, noise=None):
        """``engine.generate`` with minibatch size ``mb`` and a fresh noise seed per call.  More candidates than the
        engine's workspace holds (a population-sharded run sizes it for one shard, but the saving rank renders the
        whole population, run.py:45) go through in chunks of whole minibatches."""
        n, cap = z.shape[0], self.engine.max_population
        step = n if n <= cap else cap // mb * mb
        if step <= 0:
            raise GlassArgError(f"minibatch {mb} exceeds the engine's max_population {cap}")
        if step < n and noise is not None:
            raise GlassArgError("explicit noise tensors are not supported for chunked rendering")
        self.engine.set_batch_size(mb)
        outs = []
        for s0 in range(0, n, step):
            self._calls += 1
            seed = int(getattr(self.config, "noise_seed", 0)) + self._calls
            chunk = z if step >= n else z[s0:s0 + step].contiguous()
            outs.append(self.engine.generate(chunk, noise=noise, seed=seed))
        return outs[0] if len(outs) == 1 else torch.cat(outs)

    # generator.py:29-34
    def generate(self, ls, minibatch=None, noise=None):
        z = ls()[0]
        if self.config.task == "img2txt":
            if getattr(self.config, "return_tokens", False):          # token-level callers (no vocabulary files)
                return self.model.parse_out_tokens(self.model.generate_tokens(z))
            return self.model.generate(z, minibatch=minibatch)
        z = z.to(self.config.device, torch.float32).contiguous()
        n = z.shape[0]
        hits = self._cached_rows(z.cpu().numpy()) if noise is None else [None] * n
        missing = [i for i, h in enumerate(hits) if h is None]
        if len(missing) == n:
            return self._render(z, minibatch if minibatch is not None else n, noise)   # already normalised to [0,1]
        # some (usually all) of the requested candidates were rendered and scored by the last _evaluate
        R = self.gan.resolution
        out = torch.empty(n, 3, R, R, dtype=torch.float32, device=z.device)
        have = [i for i, h in enumerate(hits) if h is not None]
        got = self.engine.last_images([hits[i] for i in have])
        out[torch.as_tensor(have, device=z.device)] = got
        self.reuse_stats["reused"] += len(have)
        if missing:
            # render the rest: whole minibatches (models.py:112), padded by repeating the last missing row
            mb = minibatch if minibatch is not None else len(missing)
            idx = missing + [missing[-1]] * ((-len(missing)) % mb)
            extra = self._render(z[torch.as_tensor(idx, device=z.device)].contiguous(), mb)
            out[torch.as_tensor(missing, device=z.device)] = extra[:len(missing)]
            self.reuse_stats["rendered"] += len(missing)
        return out

    # generator.py:36-38
    def discriminate(self, images, minibatch=None):
        self.engine.set_batch_size(minibatch if minibatch is not None else images.shape[0])
        return self.engine.discriminate(images.contiguous())

    def has_discriminator(self):
        return False if self.config.task == "img2txt" else self.engine.use_discriminator

    # generator.py:43-51
    def clip_similarity(self, input):
        if self.config.task == "img2txt":
            return self._text_similarity(input)
        if self.augmentation is not None:
            raise NotImplementedError("augmentation hook is None in the reference (generator.py:14)")
        return self.engine.clip_similarity(input.contiguous())

    # generator.py:63-68
    def save(self, input, path):
        """utils.py:5-7 ``save_grid`` (torchvision make_grid + save_image) for P > 1, ``save_image(input[0])`` for one
        image.  Device tensors go through ``glass_image_grid_u8``: grid assembly, the x255 + 0.5 clamp, the uint8 cast
        and the CHW->HWC permute happen in one kernel and a quarter of the bytes cross PCIe; PIL encodes the file
        exactly as torchvision's save_image would (``Image.fromarray(ndarr).save(path)``)."""
        if self.config.task == "img2txt":                              # generator.py:69-71
            with open(path, "w") as f:
                f.write("\n".join(str(t) for t in input))
            return
        if isinstance(input, torch.Tensor) and input.is_cuda:
            from PIL import Image
            grid = self.engine.image_grid_u8(input.detach().float().contiguous(),
                                             nrow=8, padding=2 if input.shape[0] > 1 else 0)
            Image.fromarray(grid).save(path)
            return
        from torchvision.utils import save_image
        from .utils import save_grid
        if input.shape[0] > 1:
            save_grid(input.detach().cpu(), path)
        else:
            save_image(input[0], path)
