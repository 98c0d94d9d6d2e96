"""The UNMODIFIED reference modules (stylegan2.models / stylegan2.modules / clip.model) wrapped in the ~40 lines of
problem.py:14-29 / generator.py:29-51 / models.py:108-130 / utils.py:14-21 that cannot be imported here (pymoo /
kornia / pytorch_pretrained_biggan are not installed).  TEST / BASELINE INFRASTRUCTURE ONLY — never on the product
path.  Two users:

  * oracle/make_golden.py (build container): root = /root/reference, writes tests/golden/*.npz;
  * bench.py --impl reference (GPU box): root = oracle/_ref, a git-ignored copy of the four reference files made by
    oracle/build_ref.py in the build container (it travels with the gpurun snapshot; /root/reference does not).
    The arm then times the reference's own nets on the host cores (cpu_baseline.kind = "reference").
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

from clip_glass_b200 import weights as W

_ROOT = None
_MODS = None
HERE = os.path.dirname(os.path.abspath(__file__))
VENDORED_ROOT = os.path.join(HERE, "_ref")


def use_reference_root(root: str) -> None:
    global _ROOT, _MODS
    _ROOT, _MODS = root, None


def reference_available(root: str = None) -> bool:
    root = root or _ROOT or VENDORED_ROOT
    return os.path.exists(os.path.join(root, "stylegan2", "models.py")) and \
        os.path.exists(os.path.join(root, "clip", "model.py"))


def _import_reference():
    """(stylegan2.models, clip.model) imported from the configured root."""
    global _MODS
    if _MODS is None:
        root = _ROOT or VENDORED_ROOT
        if not reference_available(root):
            raise ImportError(f"no reference modules under {root}")
        sys.path.insert(0, root)
        try:
            for name in ("stylegan2", "clip"):
                for k in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
                    del sys.modules[k]
            models = importlib.import_module("stylegan2.models")
            clip_model = importlib.import_module("clip.model")
        finally:
            sys.path.remove(root)
        _MODS = (models, clip_model)
    return _MODS


def build_reference_gan(spec: W.GanSpec, g_sd, d_sd):
    models = _import_reference()[0]
    ch = list(spec.channels)
    G = models.Generator(
        G_mapping=models.GeneratorMapping(latent_size=spec.latent_size,
                                          num_layers=spec.mapping_layers, lr_mul=0.01),
        G_synthesis=models.GeneratorSynthesis(channels=ch, latent_size=spec.latent_size))
    D = models.Discriminator(channels=ch, mbstd_group_size=spec.mbstd_group_size)
    for net, sd in ((G, g_sd), (D, d_sd)):
        full = net.state_dict()
        missing = [k for k in sd if k not in full]
        assert not missing, missing
        learnable = {k for k, _ in net.named_parameters()}
        assert learnable <= set(sd), sorted(learnable - set(sd))[:5]
        full.update(sd)
        net.load_state_dict(full)
        net.eval()
    return G, D


def build_reference_clip(spec: W.ClipSpec, visual_sd):
    _, clip_model_mod = _import_reference()
    CLIP, convert_weights = clip_model_mod.CLIP, clip_model_mod.convert_weights
    model = CLIP(spec.embed_dim, spec.resolution, spec.layers, spec.width, spec.patch,
                 77, 64, 64, 1, 1)
    with torch.no_grad():                      # clip/model.py:286,289 leaves these empty
        model.positional_embedding.normal_(std=0.01)
        model.text_projection.normal_(std=0.1)
    convert_weights(model)                     # clip/model.py:397
    built = W.clip_as_built(visual_sd)
    model.visual.load_state_dict(built)        # strict
    for k, v in model.visual.state_dict().items():
        assert v.dtype == built[k].dtype, (k, v.dtype, built[k].dtype)
    return model.eval()


def reference_evaluate(x, G, D, clip_model, text_features, batch_size, use_d, noise):
    """problem.py:14-29 over the imported reference nets."""
    out = {}
    with torch.no_grad():
        z = torch.tensor(x.astype(float)).float()                       # latent.py:38
        assert z.shape[0] % batch_size == 0                             # models.py:112
        imgs = []
        for g in range(z.shape[0] // batch_size):                       # models.py:114-116
            if noise is not None:
                G.static_noise(noise_tensors=[t.clone() for t in noise[g]])
            imgs.append(G(z[g * batch_size:(g + 1) * batch_size]))
        generated = torch.cat(imgs)
        generated = ((generated + 1) / 2.0).clip(0, 1)                  # utils.py:14-17
        image = F.interpolate(generated, size=(224, 224), mode="bilinear",
                              align_corners=False)                      # generator.py:45 (kornia stand-in)
        feats = clip_model.encode_image(image)                          # generator.py:49
        sim = torch.cosine_similarity(feats, text_features)            # generator.py:51
        out["images"] = generated
        out["features"] = feats
        out["sim"] = sim
        if use_d:
            den = generated * 2 - 1                                     # utils.py:19-21
            ds = []
            for g in range(z.shape[0] // batch_size):                   # models.py:126-128
                ds.append(D(den[g * batch_size:(g + 1) * batch_size]))
            dis = torch.cat(ds)
            out["dis"] = dis
            hinge = torch.relu(1 - dis).squeeze(1)                      # problem.py:23-24
            out["F"] = np.column_stack((-sim.cpu().numpy(), hinge.cpu().numpy()))
        else:
            out["F"] = -sim.cpu().numpy()
    return out


