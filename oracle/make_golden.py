"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference; the GPU box does not
have it):

    python -m oracle.make_golden            # tiny + full fixtures
    python -m oracle.make_golden --only tiny

For each fixture the script
  1. draws seeded weights / latents / noise with ``clip_glass_b200.weights``
     (reference state_dict key layout),
  2. loads them into the reference's own ``stylegan2.models.Generator`` /
     ``Discriminator`` and ``clip.model.CLIP`` (with ``convert_weights``),
  3. evaluates the fitness path with those modules, restating only the ~40
     lines of problem.py:14-29 / generator.py:29-51 / models.py:108-130 that
     cannot be imported here (pymoo / kornia / pytorch_pretrained_biggan are
     not installed),
  4. checks the oracle restatement against the reference outputs and
  5. writes the inputs that are not re-derivable from seeds (text features)
     and the reference outputs to tests/golden/.

The fixtures are what pins the oracle (tests/test_oracle.py) and, through it,
the CUDA path.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, REPO)

from clip_glass_b200 import weights as W          # noqa: E402
from oracle import clip_oracle, evaluate_oracle, stylegan2_oracle   # noqa: E402
from oracle.reference_modules import (build_reference_clip, build_reference_gan,   # noqa: E402
                                      reference_evaluate, use_reference_root)

use_reference_root(REF)


def make_text_features(feats: torch.Tensor, seed: int) -> torch.Tensor:
    """A [1,E] fp16 'cached text embedding'.  Random-weight CLIP features of
    different images share a large common component, so a random vector gives
    cosines of ~0 +- 0.04 (a relative tolerance is then meaningless) and the
    mean feature gives ~0.99 for every candidate (nothing would be tested).
    Mix 0.3 of the common direction with a direction inside the span of the
    candidate-specific parts: cosines land in the 0.2-0.45 range real prompts
    give, and differ between candidates."""
    g = torch.Generator().manual_seed(seed)
    f = feats.float()
    centre = f.mean(0)
    delta = f - centre
    mix = torch.randn(f.shape[0], generator=g)
    u = (mix[:, None] * delta).sum(0)
    t = 0.3 * centre / centre.norm() + 0.95 * u / u.norm()
    return (t * 10.0)[None].half()


FIXTURES = {      # must match tests/fixtures.py CONFIGS
    "tiny": dict(gan=W.TINY_GAN, clip=W.TINY_CLIP, pop=8, batch=4, seed=100),
    "full": dict(gan=W.FFHQ, clip=W.VIT_B32, pop=4, batch=4, seed=200),
    "full_b": dict(gan=W.FFHQ, clip=W.VIT_B32, pop=4, batch=4, seed=300),
    "full_c": dict(gan=W.FFHQ, clip=W.VIT_B32, pop=8, batch=4, seed=400),
    "tiny_stress": dict(gan=W.TINY_GAN, clip=W.TINY_CLIP, pop=8, batch=4, seed=500, stress=True),
    "full_stress": dict(gan=W.FFHQ, clip=W.VIT_B32, pop=4, batch=4, seed=600, stress=True),
}


def make_fixture(name: str, out_dir: str):
    cfg = FIXTURES[name]
    gan, clipspec, P, B, seed = cfg["gan"], cfg["clip"], cfg["pop"], cfg["batch"], cfg["seed"]
    t0 = time.time()
    g_sd = W.make_generator_weights(gan, seed + 0, stress=cfg.get("stress", False))
    d_sd = W.make_discriminator_weights(gan, seed + 1)
    c_sd = W.make_clip_visual_weights(clipspec, seed + 2)
    noise = W.make_noise(gan, P // B, seed + 3)
    x = W.make_latents(P, gan.latent_size, seed + 4)
    G, D = build_reference_gan(gan, g_sd, d_sd)
    clip_model = build_reference_clip(clipspec, c_sd)

    # pass 1: features only, to place the text embedding
    probe = reference_evaluate(x, G, D, clip_model, torch.zeros(1, clipspec.embed_dim).half(),
                               B, False, noise)
    text = make_text_features(probe["features"], seed + 5)
    ref = reference_evaluate(x, G, D, clip_model, text, B, True, noise)
    ref_nod = reference_evaluate(x, G, D, clip_model, text, B, False, noise)
    print(f"[{name}] reference done in {time.time() - t0:.1f}s; sim={ref['sim'].float().numpy()}")

    # the oracle restatement must reproduce the reference
    c_built = W.clip_as_built(c_sd)
    o = evaluate_oracle.evaluate(x, g_sd, d_sd, c_built, text, gan, clipspec, B, True,
                                 noise=noise, return_images=True)
    img_err = (o["images"] - ref["images"]).abs().max().item()
    f_err = np.abs(o["F"].astype(np.float64) - ref["F"].astype(np.float64)).max()
    print(f"[{name}] oracle vs reference: max|dimg|={img_err:.3e} max|dF|={f_err:.3e}")
    assert img_err < 2e-4, img_err
    assert f_err < 2e-3, f_err
    o32 = evaluate_oracle.evaluate(x, g_sd, d_sd, c_built, text, gan, clipspec, B, True,
                                   noise=noise, clip_mode="fp32")

    small = F.avg_pool2d(ref["images"], max(1, gan.resolution // 64))
    np.savez_compressed(
        os.path.join(out_dir, f"{name}.npz"),
        pop=P, batch=B, seed=seed,
        text_features=text.numpy(),
        F=ref["F"].astype(np.float32),
        F_nod=ref_nod["F"].astype(np.float32),
        sim_fp16=ref["sim"].numpy(),
        features=ref["features"].float().numpy(),
        dis=ref["dis"].numpy(),
        images_64=small.numpy().astype(np.float32),
        image_mean=ref["images"].mean(dim=(1, 2, 3)).numpy(),
        image_sq=(ref["images"] ** 2).mean(dim=(1, 2, 3)).numpy(),
        sim_oracle_fp32=-o32["F"][:, 0].astype(np.float32),
    )
    print(f"[{name}] wrote fixture ({time.time() - t0:.1f}s)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--out", default=os.path.join(REPO, "tests", "golden"))
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    torch.manual_seed(0)
    for name in FIXTURES:
        if args.only and name != args.only:
            continue
        make_fixture(name, args.out)


if __name__ == "__main__":
    main()
