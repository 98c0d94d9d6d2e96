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

def _biggan(res: int, pop: int, batch: int):                  # config.py:27-71
    return dict(
        task="txt2img", dim_z=128, num_classes=1000, latent=DeepMindBigGANLatentSpace, model="DeepMindBigGAN",
        weights=f"biggan-deep-{res}", use_discriminator=False, algorithm="ga", norm=biggan_norm, denorm=biggan_denorm,
        truncation=1.0, pop_size=pop, batch_size=batch,
        problem_args=dict(n_var=128 + 1000, n_obj=1, n_constr=128, xl=-2, xu=2),
    )


configs.update(
    # config.py:5-25 (img2txt: GPT-2 token latents scored by CLIP's text tower against the target image)
    GPT2=dict(
        task="img2txt", dim_z=20, max_tokens_len=30, max_text_len=50, encoder_size=50257, latent=GPT2LatentSpace,
        model="GPT2", use_discriminator=False, init_text="the picture of",
        weights="./gpt2/weights/gpt2-pytorch_model.bin", encoder="./gpt2/weights/encoder.json",
        vocab="./gpt2/weights/vocab.bpe", stochastic=False, algorithm="ga", pop_size=100, batch_size=25,
        problem_args=dict(n_var=20, n_obj=1, n_constr=20, xl=0, xu=50256),
    ),
    # latent arithmetic + operators are built (latent.py, operators.py); the generator is not: the reference takes it
    # from the un-vendored pytorch_pretrained_biggan package (models.py:4,69)
    DeepMindBigGAN256=_biggan(256, 64, 32),
    DeepMindBigGAN512=_biggan(512, 32, 8),
)


def get_config(name):
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
