"""Seeded synthetic weights in the reference's own state_dict layouts.

There are no pretrained checkpoints on the build or GPU boxes (no network), so
every test, fixture and benchmark uses weights drawn here from a seeded CPU
generator.  The key names and tensor shapes are exactly those produced by the
reference's modules, so that the same dicts load into

  * ``stylegan2.models.Generator`` / ``Discriminator``
    (key layout fixed by /root/reference/stylegan2/convert_from_tf.py:177-186,
    220-228, 271-274 and the module tree in stylegan2/models.py:771-896,
    1043-1191), and
  * ``clip.model.CLIP`` via ``build_model`` (clip/model.py:363-399)

and real ``G.pth`` / ``D.pth`` / ``ViT-B-32.pt`` state dicts can be fed to the
engine through the same code path (``engine.GlassEngine``).

Noise strengths and biases are zero at the reference's init
(stylegan2/modules.py:326, :266-271); here they are drawn non-zero so that the
noise / bias code paths are actually exercised by the parity tests.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List

import torch

# ffhq-config-f: channels listed last layer -> first layer, as the reference's
# Generator takes them (stylegan2/models.py:652-657); 4x4 ... 1024x1024.
FFHQ_CHANNELS = [32, 64, 128, 256, 512, 512, 512, 512, 512]


@dataclass(frozen=True)
class GanSpec:
    """Architecture hyper-parameters of one StyleGAN2 G/D pair."""
    channels: tuple = tuple(FFHQ_CHANNELS)   # last layer -> first layer
    latent_size: int = 512
    mapping_layers: int = 8
    mbstd_group_size: int = 4

    @property
    def num_blocks(self) -> int:
        return len(self.channels)

    @property
    def resolution(self) -> int:
        return 4 * 2 ** (len(self.channels) - 1)

    @property
    def num_style_layers(self) -> int:      # stylegan2/models.py:890-896
        return 2 * len(self.channels)

    @property
    def num_noise_layers(self) -> int:      # one per modulated 3x3 conv
        return 2 * len(self.channels) - 1

    def noise_shapes(self) -> List[int]:
        """Side length of each of the noise layers, in forward order."""
        out = [4]
        for i in range(1, len(self.channels)):
            out += [4 * 2 ** i, 4 * 2 ** i]
        return out


@dataclass(frozen=True)
class ClipSpec:
    """CLIP visual tower hyper-parameters (clip/model.py:201-216)."""
    width: int = 768
    layers: int = 12
    patch: int = 32
    resolution: int = 224
    embed_dim: int = 512

    @property
    def heads(self) -> int:                 # clip/model.py:267
        return self.width // 64

    @property
    def tokens(self) -> int:
        return (self.resolution // self.patch) ** 2 + 1


FFHQ = GanSpec()
VIT_B32 = ClipSpec()
# Reduced shapes used by the fast tests (same code paths, seconds on CPU).
TINY_GAN = GanSpec(channels=(32, 32, 64, 64, 64))          # 64x64 images
TINY_CLIP = ClipSpec(width=128, layers=2)


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def make_generator_weights(spec: GanSpec = FFHQ, seed: int = 0, stress: bool = False) -> Dict[str, torch.Tensor]:
    """State dict (learnable tensors only) for the reference ``Generator``.

    ``stress``: fp16-range stress variant (the reference runs G in fp32).  Trained ffhq styles reach magnitudes far
    above the ``1 + N(0, 0.1^2)`` drawn here, so the style-affine biases of the modulated 3x3 convs are scaled by
    30 on four input channels (even layers) or by 20000 on two (odd layers: activation x style then exceeds the
    fp16 maximum of 65504 unless the style is normalised), and three output channels of every conv weight by 10."""
    g = _gen(seed)
    n = lambda *s, std=1.0: torch.randn(*s, generator=g) * std
    L = spec.latent_size
    sd: Dict[str, torch.Tensor] = {}
    for i in range(spec.mapping_layers):
        # lr_mul 0.01 => init std 1/lr_mul (stylegan2/modules.py:106-108)
        sd[f"G_mapping.main.{i}.layer.weight"] = n(L, L, std=100.0)
        sd[f"G_mapping.main.{i}.bias"] = n(L, std=10.0)
    ch = list(spec.channels)[::-1]           # first layer -> last layer
    sd["G_synthesis.const"] = n(ch[0], 4, 4)

    def mod_conv(prefix, cout, cin, k):
        sd[prefix + ".weight"] = n(cout, cin, k, k)
        sd[prefix + ".dense.layer.weight"] = n(cin, L)
        sd[prefix + ".dense.bias"] = 1.0 + n(cin, std=0.1)

    for b in range(spec.num_blocks):
        cin = ch[max(b - 1, 0)]
        cout = ch[b]
        nl = 1 if b == 0 else 2
        for l in range(nl):
            p = f"G_synthesis.conv_blocks.{b}.conv_block.{l}"
            mod_conv(p + ".layer.layer", cout, cin if l == 0 else cout, 3)
            sd[p + ".layer.weight"] = n(1, std=0.1)          # noise strength
            sd[p + ".bias"] = n(cout, std=0.1)
        p = f"G_synthesis.to_data_layers.{b}"
        mod_conv(p + ".layer", 3, cout, 1)
        # keep the summed RGB mostly inside [-1,1] so that the [0,1] clip of
        # utils.py:14-17 does not hide errors behind saturation
        sd[p + ".layer.weight"] *= 0.15
        sd[p + ".bias"] = n(3, std=0.1)
    if stress:
        li = 0
        for b in range(spec.num_blocks):
            for l in range(1 if b == 0 else 2):
                p = f"G_synthesis.conv_blocks.{b}.conv_block.{l}.layer.layer"
                if li % 2 == 0:
                    sd[p + ".dense.bias"][0:4] *= 30.0
                else:
                    sd[p + ".dense.bias"][4:6] *= 20000.0
                sd[p + ".weight"][0:3] *= 10.0
                li += 1
    return sd


def make_discriminator_weights(spec: GanSpec = FFHQ, seed: int = 1) -> Dict[str, torch.Tensor]:
    """State dict (learnable tensors only) for the reference ``Discriminator``."""
    g = _gen(seed)
    n = lambda *s, std=1.0: torch.randn(*s, generator=g) * std
    ch = list(spec.channels)                 # D goes first layer -> last layer
    sd: Dict[str, torch.Tensor] = {}
    sd["from_data_layers.0.layer.weight"] = n(ch[0], 3, 1, 1)
    sd["from_data_layers.0.bias"] = n(ch[0], std=0.1)
    for b in range(len(ch) - 1):
        p = f"conv_blocks.{b}"
        sd[p + ".conv_block.0.layer.weight"] = n(ch[b], ch[b], 3, 3)
        sd[p + ".conv_block.0.bias"] = n(ch[b], std=0.1)
        sd[p + ".conv_block.1.layer.weight"] = n(ch[b + 1], ch[b], 3, 3)
        sd[p + ".conv_block.1.bias"] = n(ch[b + 1], std=0.1)
        sd[p + ".projection.weight"] = n(ch[b + 1], ch[b], 1, 1)
    last = len(ch) - 1
    extra = 1 if spec.mbstd_group_size else 0
    sd[f"conv_blocks.{last}.1.conv_block.0.layer.weight"] = n(ch[-1], ch[-1] + extra, 3, 3)
    sd[f"conv_blocks.{last}.1.conv_block.0.bias"] = n(ch[-1], std=0.1)
    sd["dense.0.layer.weight"] = n(ch[-1], ch[-1] * 16)
    sd["dense.0.bias"] = n(ch[-1], std=0.1)
    sd["dense.1.layer.weight"] = n(1, ch[-1])
    sd["dense.1.bias"] = n(1, std=0.1)
    return sd


def make_clip_visual_weights(spec: ClipSpec = VIT_B32, seed: int = 2) -> Dict[str, torch.Tensor]:
    """fp32 master weights for ``CLIP.visual`` with the reference's key names
    (``visual.*`` stripped).  ``clip_as_built`` applies the fp16 conversion the
    reference does in ``convert_weights`` (clip/model.py:339-360)."""
    g = _gen(seed)
    n = lambda *s, std=1.0: torch.randn(*s, generator=g) * std
    W, Ly = spec.width, spec.layers
    sd: Dict[str, torch.Tensor] = {}
    sd["conv1.weight"] = n(W, 3, spec.patch, spec.patch, std=(3 * spec.patch ** 2) ** -0.5)
    sd["class_embedding"] = n(W, std=W ** -0.5)
    sd["positional_embedding"] = n(spec.tokens, W, std=W ** -0.5)
    sd["ln_pre.weight"] = 1.0 + n(W, std=0.1)
    sd["ln_pre.bias"] = n(W, std=0.1)
    attn_std = W ** -0.5
    proj_std = (W ** -0.5) * ((2 * Ly) ** -0.5)
    fc_std = (2 * W) ** -0.5
    for l in range(Ly):
        p = f"transformer.resblocks.{l}"
        sd[p + ".attn.in_proj_weight"] = n(3 * W, W, std=attn_std)
        sd[p + ".attn.in_proj_bias"] = n(3 * W, std=0.02)
        sd[p + ".attn.out_proj.weight"] = n(W, W, std=proj_std)
        sd[p + ".attn.out_proj.bias"] = n(W, std=0.02)
        sd[p + ".ln_1.weight"] = 1.0 + n(W, std=0.1)
        sd[p + ".ln_1.bias"] = n(W, std=0.1)
        sd[p + ".mlp.c_fc.weight"] = n(4 * W, W, std=fc_std)
        sd[p + ".mlp.c_fc.bias"] = n(4 * W, std=0.02)
        sd[p + ".mlp.c_proj.weight"] = n(W, 4 * W, std=proj_std)
        sd[p + ".mlp.c_proj.bias"] = n(W, std=0.02)
        sd[p + ".ln_2.weight"] = 1.0 + n(W, std=0.1)
        sd[p + ".ln_2.bias"] = n(W, std=0.1)
    sd["ln_post.weight"] = 1.0 + n(W, std=0.1)
    sd["ln_post.bias"] = n(W, std=0.1)
    sd["proj"] = n(W, spec.embed_dim, std=W ** -0.5)
    return sd


_CLIP_FP32_KEYS = ("class_embedding", "positional_embedding", "ln_")


def clip_as_built(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Apply ``convert_weights`` (clip/model.py:339-360): Conv/Linear/MHA
    parameters and ``proj`` become fp16; LayerNorm, class and positional
    embeddings stay fp32."""
    out = {}
    for k, v in sd.items():
        keep32 = any(t in k for t in _CLIP_FP32_KEYS)
        out[k] = v.float() if keep32 else v.half()
    return out


def make_noise(spec: GanSpec, n_groups: int, seed: int = 3) -> List[List[torch.Tensor]]:
    """Explicit noise tensors: one list of ``num_noise_layers`` fp32 tensors of
    shape [1,1,H,W] per minibatch group, i.e. what the reference would draw
    with ``normal_()`` once per ``G(z_minibatch)`` call
    (stylegan2/modules.py:426-452, models.py:114-116)."""
    g = _gen(seed)
    return [[torch.randn(1, 1, s, s, generator=g) for s in spec.noise_shapes()]
            for _ in range(n_groups)]


def make_latents(pop: int, dim: int = 512, seed: int = 4):
    """Population as pymoo hands it to ``_evaluate``: float64 [P, n_var]
    (operators.py:24-25 draws N(0,1); config.py:91-92 bounds to [-10,10])."""
    import numpy as np
    x = np.random.default_rng(seed).normal(0.0, 1.0, size=(pop, dim))
    return np.clip(x, -10.0, 10.0)


# ---------------------------------------------------------------------------
# real checkpoints: the reference's pickle-dict format (stylegan2/models.py:111-132, 160-196, 249-262)
# ---------------------------------------------------------------------------
def flatten_reference_blob(blob, prefix: str = ""):
    """``G.pth`` / ``D.pth`` as written by the reference's ``_serialize``: a dict
    ``{'name', 'kwargs', 'state_dict', <sub-model name>: <the same, recursively>}``.  ``Generator`` overrides
    ``_get_state_dict`` (stylegan2/models.py:249-262), so its top-level ``state_dict`` holds only the ``dlatent_avg``
    buffer and the weights sit under ``blob['G_mapping']['state_dict']`` (keys ``main.0...``) and
    ``blob['G_synthesis']['state_dict']`` (keys ``const``, ``conv_blocks...``).  Returns
    ``(flat_state_dict, kwargs_by_prefix)`` with the sub-model names as key prefixes — the layout
    ``nn.Module.state_dict()`` of the assembled model has and ``packing.pack_*`` read."""
    if not (isinstance(blob, dict) and "state_dict" in blob):
        return dict(blob), {prefix.rstrip("."): {}}            # a bare state_dict
    flat = {prefix + k: v for k, v in blob["state_dict"].items()}
    kwargs = {prefix.rstrip("."): dict(blob.get("kwargs", {}))}
    for key, sub in blob.items():
        if key in ("name", "kwargs", "state_dict"):
            continue
        if isinstance(sub, dict) and "state_dict" in sub:
            f2, k2 = flatten_reference_blob(sub, prefix + key + ".")
            flat.update(f2)
            kwargs.update(k2)
    return flat, kwargs


def load_reference_checkpoint(path: str):
    """torch.load + ``flatten_reference_blob``."""
    blob = torch.load(path, map_location="cpu", weights_only=False)
    return flatten_reference_blob(blob)


class UnsupportedArchitecture(ValueError):
    pass


def gan_spec_from_checkpoint(g_sd, g_kwargs, d_kwargs=None) -> GanSpec:
    """Architecture hyper-parameters from the checkpoint itself (the reference rebuilds the nets from the pickled
    kwargs, stylegan2/models.py:160-180; config.py only names the weight folder), so that the church / car / ffhq
    config-f checkpoints all load without a hand-written spec.  Shapes are read from the tensors, flags from kwargs;
    anything the CUDA path does not implement fails here with a message instead of a KeyError in packing."""
    syn = g_kwargs.get("G_synthesis", {})
    mp = g_kwargs.get("G_mapping", {})
    n_blocks = 0
    while f"G_synthesis.to_data_layers.{n_blocks}.layer.weight" in g_sd:
        n_blocks += 1
    if n_blocks < 2:
        raise UnsupportedArchitecture("no G_synthesis.to_data_layers.* tensors: not a skip-architecture StyleGAN2 generator")
    ch_first_to_last = [int(g_sd[f"G_synthesis.to_data_layers.{b}.layer.weight"].shape[1]) for b in range(n_blocks)]
    n_map = 0
    while f"G_mapping.main.{n_map}.layer.weight" in g_sd:
        n_map += 1
    latent = int(g_sd["G_mapping.main.0.layer.weight"].shape[1])
    problems = []
    if syn.get("resnet", False) or not syn.get("skip", True):
        problems.append("G_synthesis must be the skip architecture (skip=True, resnet=False)")
    if syn.get("data_channels", 3) != 3:
        problems.append("data_channels must be 3")
    if mp.get("label_size", 0):
        problems.append("conditional mapping networks (label_size > 0) are not supported")
    if any(c % 32 or c <= 0 or c > 512 for c in ch_first_to_last):
        problems.append(f"channels {ch_first_to_last} must be multiples of 32 in [32, 512]")
    if tuple(g_sd["G_synthesis.const"].shape[1:]) != (4, 4):
        problems.append("base resolution must be 4x4")
    for key in ("conv_filter", "skip_filter", "conv_resample_filter"):
        f = syn.get(key, [1, 3, 3, 1])
        if list(f) != [1, 3, 3, 1]:
            problems.append(f"{key} must be [1,3,3,1]")
    if str(syn.get("activation", "leaky:0.2")) not in ("leaky:0.2", "lrelu:0.2"):
        problems.append("activation must be leaky:0.2")
    d_kwargs = (d_kwargs or {}).get("", d_kwargs or {})
    if d_kwargs:
        if not d_kwargs.get("resnet", True) or d_kwargs.get("skip", False):
            problems.append("the discriminator must be the resnet architecture")
        if d_kwargs.get("mbstd_group_size", 4) not in (0, 1, 2, 4, 8) and d_kwargs.get("mbstd_group_size") is not None:
            problems.append("mbstd_group_size must be <= 8")
    if problems:
        raise UnsupportedArchitecture("checkpoint architecture not supported by the B200 path: " + "; ".join(problems))
    mb = d_kwargs.get("mbstd_group_size", 4) if d_kwargs else 4
    return GanSpec(channels=tuple(ch_first_to_last[::-1]), latent_size=latent, mapping_layers=n_map,
                   mbstd_group_size=int(mb or 0))
