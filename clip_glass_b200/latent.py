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


class DeepMindBigGANLatentSpace(torch.nn.Module):
    """latent.py:4-24.  ``forward() -> (clip(z, -2, 2), softmax(class_labels, dim=1))``.  With the population set by
    ``set_from_population`` and a CUDA device the arithmetic runs in ``glass_biggan_latent`` (one kernel: f64 -> f32,
    clip, row softmax over the 1000 class genes); values set through ``set_values`` take the torch ops the reference
    uses.  The BigGAN generator itself (pytorch_pretrained_biggan 0.1.1) is not vendored by the reference and is not
    part of this build."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.z = torch.nn.Parameter(torch.zeros(0, config.dim_z), requires_grad=False)
        self.class_labels = torch.nn.Parameter(torch.zeros(0, config.num_classes), requires_grad=False)
        self.population = None

    def set_values(self, z, class_labels):
        self.z.data = z
        self.class_labels.data = class_labels
        self.population = None

    def set_from_population(self, x):
        # latent.py:16-18
        self.population = np.ascontiguousarray(x.astype(float))
        dz = self.config.dim_z
        self.z.data = torch.from_numpy(self.population[:, :dz]).float().to(self.config.device)
        self.class_labels.data = torch.from_numpy(self.population[:, dz:]).float().to(self.config.device)

    def forward(self):
        if self.population is not None and self.z.is_cuda:
            import ctypes
            from ._lib import check, load_library
            P, dz, nc = self.population.shape[0], self.config.dim_z, self.config.num_classes
            z = torch.empty(P, dz, dtype=torch.float32, device=self.z.device)
            cl = torch.empty(P, nc, dtype=torch.float32, device=self.z.device)
            with torch.cuda.device(self.z.device):
                stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
                check(load_library().glass_biggan_latent(self.population.ctypes.data, P, dz, nc, z.data_ptr(),
                                                         cl.data_ptr(), stream))
            return z, cl
        return torch.clip(self.z, -2, 2), torch.softmax(self.class_labels, dim=1)      # latent.py:20-24


class GPT2LatentSpace(torch.nn.Module):
    """latent.py:44-59: integer token latents [P, dim_z] in [0, encoder_size)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.z = torch.zeros(0, config.dim_z, dtype=torch.long)
        self.population = None

    def set_values(self, z):
        self.z = z
        self.population = None

    def set_from_population(self, x):
        # latent.py:55-56: torch.tensor(x.astype(int)).long().to(device)
        self.population = np.ascontiguousarray(x.astype(int))
        self.z = torch.from_numpy(self.population).long().to(self.config.device)

    def forward(self):
        return (self.z,)
