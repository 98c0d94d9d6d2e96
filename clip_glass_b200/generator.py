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

import torch

from . import weights as W
from .engine import GlassEngine
from ._lib import GlassError


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
        def sd_of(path):
            blob = torch.load(path, map_location="cpu", weights_only=False)
            sd = blob["state_dict"] if isinstance(blob, dict) and "state_dict" in blob else blob
            return {k: v for k, v in sd.items()}
        g_sd = sd_of(os.path.join(wdir, "G.pth"))
        d_sd = sd_of(os.path.join(wdir, "D.pth")) if config.use_discriminator else None
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
        if config.task != "txt2img":
            raise NotImplementedError("img2txt (GPT-2) is a 'next' row of SURVEY.md §8(f)")
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

    # generator.py:29-34
    def generate(self, ls, minibatch=None, noise=None):
        z = ls()[0]
        z = z.to(self.config.device, torch.float32).contiguous()
        self.engine.set_batch_size(minibatch if minibatch is not None else z.shape[0])
        self._calls += 1
        seed = int(getattr(self.config, "noise_seed", 0)) + self._calls
        return self.engine.generate(z, noise=noise, seed=seed)     # already normalised to [0,1]

    # generator.py:36-38
    def discriminate(self, images, minibatch=None):
        self.engine.set_batch_size(minibatch if minibatch is not None else images.shape[0])
        return self.engine.discriminate(images.contiguous())

    def has_discriminator(self):
        return self.engine.use_discriminator

    # generator.py:43-51
    def clip_similarity(self, input):
        if self.augmentation is not None:
            raise NotImplementedError("augmentation hook is None in the reference (generator.py:14)")
        return self.engine.clip_similarity(input.contiguous())

    # generator.py:63-68
    def save(self, input, path):
        from torchvision.utils import save_image
        from .utils import save_grid
        if input.shape[0] > 1:
            save_grid(input.detach().cpu(), path)
        else:
            save_image(input[0], path)
