"""Mirror of the reference's config.py for the StyleGAN2 configs.

Same keys and values as config.py:74-195 of the reference (task, dim_z,
latent, use_discriminator, weights, algorithm, norm/denorm, pop_size,
batch_size, problem_args); ``model`` is the B200 engine-backed StyleGAN2
stand-in.  ``get_config`` returns a *copy* so BASELINE.json's population
overrides (``pop_size`` has no CLI flag in the reference) do not leak.

Extra keys understood by this build (all optional):
  text_features   [1,512] tensor: the cached CLIP.encode_text result
                  (generator.py:23-24); required unless weights are synthetic
  synthetic_seed  int: draw seeded random weights instead of reading
                  G.pth / D.pth / ViT-B-32.pt (none exist offline)
  gan_spec / clip_spec  override the architecture (tests use reduced shapes)
  noise_seed      seed of the device noise generator (per-generation offset added)
"""
from __future__ import annotations

import copy

from .latent import DeepMindBigGANLatentSpace, GPT2LatentSpace, StyleGAN2LatentSpace
from .utils import biggan_denorm, biggan_norm


def _stylegan2(weights: str, use_d: bool):
    return dict(
        task="txt2img",
        dim_z=512,
        latent=StyleGAN2LatentSpace,
        model="StyleGAN2",
        use_discriminator=use_d,
        weights=weights,
        algorithm="nsga2" if use_d else "ga",
        norm=biggan_norm,
        denorm=biggan_denorm,
        pop_size=16,
        batch_size=4,
        problem_args=dict(n_var=512, n_obj=2 if use_d else 1, n_constr=512, xl=-10, xu=10),
    )


configs = dict(
    StyleGAN2_ffhq_d=_stylegan2("./stylegan2/weights/ffhq-config-f", True),
    StyleGAN2_car_d=_stylegan2("./stylegan2/weights/car-config-f", True),
    StyleGAN2_church_d=_stylegan2("./stylegan2/weights/church-config-f", True),
    StyleGAN2_ffhq_nod=_stylegan2("./stylegan2/weights/ffhq-config-f", False),
    StyleGAN2_car_nod=_stylegan2("./stylegan2/weights/car-config-f", False),
    StyleGAN2_church_nod=_stylegan2("./stylegan2/weights/church-config-f", False),
)

_OUT_OF_SCOPE = {"GPT2": GPT2LatentSpace, "DeepMindBigGAN256": DeepMindBigGANLatentSpace,
                 "DeepMindBigGAN512": DeepMindBigGANLatentSpace}


def get_config(name):
    if name in _OUT_OF_SCOPE:
        raise NotImplementedError(
            f"config {name!r}: the BigGAN / GPT-2 paths are 'next' rows of SURVEY.md §8(f), not built yet")
    return copy.deepcopy(configs[name])


class Namespace:
    """What run.py:24-25 builds: argparse namespace overlaid with the config dict."""
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def __contains__(self, k):
        return k in self.__dict__


def make_namespace(name: str, device: str = "cuda", target: str = "", **overrides) -> Namespace:
    ns = Namespace(device=device, config=name, generations=500, save_each=50, tmp_folder="./tmp", target=target)
    ns.__dict__.update(get_config(name))
    ns.__dict__.update(overrides)
    return ns
