"""Shared fixture builders: seeded weights / latents / noise for the two
golden configurations written by oracle/make_golden.py."""
import os

import numpy as np
import torch

from clip_glass_b200 import weights as W

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# must match oracle/make_golden.py FIXTURES
CONFIGS = {
    "tiny": dict(gan=W.TINY_GAN, clip=W.TINY_CLIP, pop=8, batch=4, seed=100),
    "full": dict(gan=W.FFHQ, clip=W.VIT_B32, pop=4, batch=4, seed=200),
}


def load_golden(name):
    return dict(np.load(os.path.join(REPO, "tests", "golden", f"{name}.npz")))


def build_inputs(name):
    cfg = CONFIGS[name]
    seed = cfg["seed"]
    gan, clip = cfg["gan"], cfg["clip"]
    return dict(
        gan=gan, clip=clip, pop=cfg["pop"], batch=cfg["batch"],
        g_sd=W.make_generator_weights(gan, seed + 0),
        d_sd=W.make_discriminator_weights(gan, seed + 1),
        c_sd=W.make_clip_visual_weights(clip, seed + 2),
        noise=W.make_noise(gan, cfg["pop"] // cfg["batch"], seed + 3),
        x=W.make_latents(cfg["pop"], gan.latent_size, seed + 4),
    )
