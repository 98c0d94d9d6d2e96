"""Mirror of the reference's latent.py for the StyleGAN2 path.

``StyleGAN2LatentSpace`` keeps the reference's surface (latent.py:27-41):
``set_values``, ``set_from_population(ndarray)``, ``forward() -> (z,)`` and a
``state_dict()`` holding ``z`` (run.py:98-101 saves it).  It does not allocate
the throw-away ``randn`` Parameter the reference creates on every
``_evaluate`` call (latent.py:32) — the values are never read before being
overwritten.
"""
from __future__ import annotations

import numpy as np
import torch


class StyleGAN2LatentSpace(torch.nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.z = torch.nn.Parameter(torch.zeros(0, config.dim_z), requires_grad=False)
        self.population = None          # the float64 ndarray, for the fused host entry point

    def set_values(self, z):
        self.z.data = z
        self.population = None

    def set_from_population(self, x):
        # latent.py:38: torch.tensor(x.astype(float)).float().to(device)
        self.population = np.ascontiguousarray(x.astype(float))
        self.z.data = torch.from_numpy(self.population).float().to(self.config.device)

    def forward(self):
        return (self.z,)


class DeepMindBigGANLatentSpace:   # latent.py:4-24 — SURVEY.md §8(f) row 3, not on the built path
    def __init__(self, config):
        raise NotImplementedError("BigGAN latent space is outside the B200 hot-path scope (SURVEY.md §8f)")


class GPT2LatentSpace:             # latent.py:44-59 — SURVEY.md §8(f) row 2
    def __init__(self, config):
        raise NotImplementedError("GPT-2 token-latent path is outside the B200 hot-path scope (SURVEY.md §8f)")
