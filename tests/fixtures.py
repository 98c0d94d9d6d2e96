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
    # round 2: more weight / latent / noise seeds at full size; "full_c" has two noise groups (P=8)
    "full_b": dict(gan=W.FFHQ, clip=W.VIT_B32, pop=4, batch=4, seed=300),
    "full_c": dict(gan=W.FFHQ, clip=W.VIT_B32, pop=8, batch=4, seed=400),
    # fp16-range stress (weights.make_generator_weights(stress=True)): large style magnitudes
    "tiny_stress": dict(gan=W.TINY_GAN, clip=W.TINY_CLIP, pop=8, batch=4, seed=500, stress=True),
    "full_stress": dict(gan=W.FFHQ, clip=W.VIT_B32, pop=4, batch=4, seed=600, stress=True),
}
FULL_FIXTURES = ("full", "full_b", "full_c", "full_stress")


def load_golden(name):
    return dict(np.load(os.path.join(REPO, "tests", "golden", f"{name}.npz")))


def build_inputs(name):
    cfg = CONFIGS[name]
    seed = cfg["seed"]
    gan, clip = cfg["gan"], cfg["clip"]
    return dict(
        gan=gan, clip=clip, pop=cfg["pop"], batch=cfg["batch"],
        g_sd=W.make_generator_weights(gan, seed + 0, stress=cfg.get("stress", False)),
        d_sd=W.make_discriminator_weights(gan, seed + 1),
        c_sd=W.make_clip_visual_weights(clip, seed + 2),
        noise=W.make_noise(gan, cfg["pop"] // cfg["batch"], seed + 3),
        x=W.make_latents(cfg["pop"], gan.latent_size, seed + 4),
    )
