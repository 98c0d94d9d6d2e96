"""ORACLE (test infrastructure only — never on the product path).

CPU restatement of CLIP-GLaSS's per-generation fitness evaluation,
``GenerationProblem._evaluate`` (problem.py:14-29) and the façade it calls
(generator.py:29-51, models.py:108-130, latent.py:37-41, utils.py:14-21),
for the StyleGAN2 txt2img configs.  The arithmetic of the nets is in
``stylegan2_oracle`` / ``clip_oracle``.

Third-party arithmetic that is NOT under /root/reference:
``kornia.resize`` (kornia==0.4.1, requirements.txt:20; call site
generator.py:45) is restated as bilinear, align_corners=False, no antialias
(``F.interpolate``) — parity with kornia itself is unpinned (kornia is not
installable here); everything else is pinned by tests/golden via
``oracle/make_golden.py``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import clip_oracle, stylegan2_oracle


def biggan_norm(x: torch.Tensor) -> torch.Tensor:
    """utils.py:14-17."""
    return ((x + 1) / 2.0).clip(0, 1)


def biggan_denorm(x: torch.Tensor) -> torch.Tensor:
    """utils.py:19-21."""
    return x * 2 - 1


def latents_from_population(x: np.ndarray) -> torch.Tensor:
    """latent.py:37-38: f64 ndarray -> fp32 tensor."""
    return torch.tensor(x.astype(float)).float()


def generate(z: torch.Tensor, g_sd, gan_spec, minibatch: int,
             noise: Optional[Sequence[Sequence[torch.Tensor]]]) -> torch.Tensor:
    """generator.py:29-34 + models.py:108-118: serial minibatch loop over G,
    cat, then ``config.norm``.  ``noise[g]`` is the per-minibatch noise list
    (the reference redraws it on every ``G(z_minibatch)`` call)."""
    assert z.shape[0] % minibatch == 0                                  # models.py:112
    out = []
    for g in range(z.shape[0] // minibatch):
        zb = z[g * minibatch:(g + 1) * minibatch]
        out.append(stylegan2_oracle.generator(
            zb, g_sd, gan_spec.num_blocks, gan_spec.mapping_layers,
            None if noise is None else noise[g]))
    return biggan_norm(torch.cat(out))


def resize224(images: torch.Tensor, size: int = 224) -> torch.Tensor:
    """generator.py:45 ``kornia.resize(input, (224, 224))``."""
    return F.interpolate(images, size=(size, size), mode="bilinear", align_corners=False)


def clip_similarity(images: torch.Tensor, clip_sd, clip_spec, text_features: torch.Tensor,
                    mode: str = "as_built") -> torch.Tensor:
    """generator.py:43-51 (txt2img branch).  NB: no CLIP mean/std
    normalisation is applied by the reference.  The whole population goes
    through CLIP as one batch."""
    feats = clip_oracle.encode_image(resize224(images, clip_spec.resolution), clip_sd,
                                     clip_spec.layers, clip_spec.heads, mode)
    return torch.cosine_similarity(feats, text_features.to(feats.dtype))


def discriminate(images: torch.Tensor, d_sd, gan_spec, minibatch: int) -> torch.Tensor:
    """generator.py:36-38 + models.py:120-130: denorm (of the CLIPPED image)
    then serial minibatch loop over D, cat."""
    images = biggan_denorm(images)
    assert images.shape[0] % minibatch == 0                             # models.py:124
    out = []
    for g in range(images.shape[0] // minibatch):
        out.append(stylegan2_oracle.discriminator(
            images[g * minibatch:(g + 1) * minibatch], d_sd, gan_spec.num_blocks,
            gan_spec.mbstd_group_size))
    return torch.cat(out)


def evaluate(x: np.ndarray, g_sd, d_sd, clip_sd, text_features: torch.Tensor,
             gan_spec, clip_spec, batch_size: int, use_discriminator: bool,
             noise=None, clip_mode: str = "as_built", return_images: bool = False):
    """problem.py:14-29.  Returns ``F`` exactly as pymoo receives it: [P]
    (``-sim``) for the single-objective configs or [P,2] ``(-sim, hinge)`` for
    the NSGA-II configs, plus ``G`` = zeros[P]."""
    with torch.no_grad():
        z = latents_from_population(x)
        generated = generate(z, g_sd, gan_spec, batch_size, noise)
        sim = clip_similarity(generated, clip_sd, clip_spec, text_features, clip_mode).cpu().numpy()
        if use_discriminator:
            dis = discriminate(generated, d_sd, gan_spec, batch_size)
            hinge = torch.relu(1 - dis).squeeze(1).cpu().numpy()
            Fv = np.column_stack((-sim, hinge))
        else:
            Fv = -sim
    out = {"F": Fv, "G": np.zeros((x.shape[0]))}
    if return_images:
        out["images"] = generated
    return out
