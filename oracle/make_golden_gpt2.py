"""Generate tests/golden/gpt2_*.npz by running the UNMODIFIED reference modules of the img2txt path
(BASELINE config 5): gpt2.model.GPT2LMHeadModel + gpt2.sample.sample_sequence (models.py:45-60) and
clip.model.CLIP.encode_text (generator.py:53-59).  Build container only (needs /root/reference):

    python -m oracle.make_golden_gpt2 [--only gpt2_tiny]

What cannot be pinned here, and why (stated in DESIGN.md too):
  * the CLIP BPE vocabulary (clip/bpe_simple_vocab_16e6.txt.gz) is NOT in the reference tree and ftfy is not installed,
    so ``clip.tokenize`` cannot run: the fixtures carry CLIP token rows built from the generated GPT-2 tokens by a
    documented stand-in map (SOT, mapped tokens, EOT, zero padding — the shape clip/clip.py:125-139 produces), which
    exercises variable lengths and the EOT-argmax gather;
  * the GPT-2 BPE decode (gpt2/encoder.py, vocabulary present) is string work on the host; it is pinned by
    tests/test_host_cpu.py against the reference's Encoder, not by these fixtures.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, REPO)

from clip_glass_b200 import text_weights as TW      # noqa: E402
from oracle import gpt2_oracle                       # noqa: E402

INIT_TOKENS = [1169, 4286, 286]                      # gpt2 BPE of config.init_text "the picture of" (config.py:15)
DIM_Z, LENGTH = 20, 30                               # config.py:8-9


def reference_modules():
    sys.path.insert(0, REF)
    try:
        from gpt2.config import GPT2Config
        from gpt2.model import GPT2LMHeadModel
        from gpt2.sample import sample_sequence
        from gpt2.utils import load_weight
        from clip.model import CLIP, convert_weights
    finally:
        sys.path.remove(REF)
    return GPT2Config, GPT2LMHeadModel, sample_sequence, load_weight, CLIP, convert_weights


def build_reference_gpt2(spec: TW.GPT2Spec, sd):
    GPT2Config, GPT2LMHeadModel, _, load_weight, _, _ = reference_modules()
    cfg = GPT2Config(vocab_size_or_config_json_file=spec.vocab, n_positions=spec.n_positions, n_ctx=spec.n_positions,
                     n_embd=spec.n_embd, n_layer=spec.n_layer, n_head=spec.n_head, layer_norm_epsilon=spec.eps)
    model = GPT2LMHeadModel(cfg)
    model = load_weight(model, {k: v.clone() for k, v in sd.items()})          # models.py:26-27
    return model.eval()


def build_reference_clip_text(spec: TW.ClipTextSpec, sd):
    _, _, _, _, CLIP, convert_weights = reference_modules()
    model = CLIP(spec.embed_dim, 64, 1, 64, 32, spec.context, spec.vocab, spec.width, spec.heads, spec.layers)
    with torch.no_grad():
        model.visual.positional_embedding.normal_(std=0.01)
    convert_weights(model)
    built = TW.text_as_built(sd)
    missing, unexpected = model.load_state_dict(built, strict=False)
    assert not unexpected and all(k.startswith("visual.") or k == "logit_scale" for k in missing), (missing, unexpected)
    for k, v in built.items():
        assert model.state_dict()[k].dtype == v.dtype, (k, model.state_dict()[k].dtype, v.dtype)
    return model.eval()


def standin_clip_tokens(gen_tokens, spec: TW.ClipTextSpec, rng) -> np.ndarray:
    """[P, context] int64 rows shaped like clip.tokenize output (clip/clip.py:125-139): SOT, body, EOT, zeros.  The
    body is the generated GPT-2 token list cut to a per-row length and mapped into CLIP's id range below SOT."""
    sot, eot = spec.vocab - 2, spec.vocab - 1
    out = np.zeros((len(gen_tokens), spec.context), dtype=np.int64)
    for i, toks in enumerate(gen_tokens):
        n = int(rng.integers(1, min(len(toks), spec.context - 2) + 1)) if len(toks) else 0
        body = [(int(t) * 7 + 13) % (spec.vocab - 258) + 256 for t in toks[:n]]
        row = [sot] + body + [eot]
        out[i, :len(row)] = row
    return out


FIXTURES = {
    "gpt2_tiny": dict(gpt2=TW.TINY_GPT2, text=TW.TINY_CLIP_TEXT, pop=8, seed=700),
    "gpt2_full": dict(gpt2=TW.GPT2_SMALL, text=TW.CLIP_TEXT_B32, pop=8, seed=800),
}


def make_fixture(name: str, out_dir: str):
    cfg = FIXTURES[name]
    gspec, tspec, P, seed = cfg["gpt2"], cfg["text"], cfg["pop"], cfg["seed"]
    t0 = time.time()
    g_sd = TW.make_gpt2_weights(gspec, seed)
    t_sd = TW.make_clip_text_weights(tspec, seed + 1)
    z = TW.make_token_latents(P, DIM_Z, gspec.vocab, seed + 2)
    eot = gspec.vocab - 1
    z[1, 7] = eot                       # an EOT gene inside the latent part: parse_out then yields the empty text
    init = [t % gspec.vocab for t in INIT_TOKENS]
    _, _, sample_sequence, _, _, _ = reference_modules()
    model = build_reference_gpt2(gspec, g_sd)
    ctx = torch.cat((torch.tensor(z).long(), torch.tensor(init).long().repeat(P, 1)), dim=1)       # models.py:47-48
    ref_tokens = np.asarray(sample_sequence(model=model, length=LENGTH, context=ctx, start_token=None, batch_size=P,
                                            temperature=0.7, top_k=40, device="cpu", sample=False), dtype=np.int64)
    print(f"[{name}] reference GPT-2 decode done in {time.time() - t0:.1f}s")
    ora_tokens, first_logits = gpt2_oracle.gpt2_generate_tokens(g_sd, gspec, z, init, LENGTH, return_logits=True)
    assert np.array_equal(ora_tokens, ref_tokens), "oracle GPT-2 tokens differ from the reference"
    top2 = np.sort(first_logits, axis=1)[:, -2:]
    gap = (top2[:, 1] - top2[:, 0])
    print(f"[{name}] tokens bit-exact; first-step top-1/top-2 logit gap: min {gap.min():.3e}, logit std {first_logits.std():.3f}")
    gen = gpt2_oracle.parse_out_tokens(ref_tokens, DIM_Z, eot)
    assert gen[1] == []                 # the planted EOT gene
    rng = np.random.default_rng(seed + 3)
    clip_tokens = standin_clip_tokens(gen, tspec, rng)
    clip_model = build_reference_clip_text(tspec, t_sd)
    with torch.no_grad():
        ref_feat = clip_model.encode_text(torch.tensor(clip_tokens))                                # generator.py:57
    built = TW.text_as_built(t_sd)
    ora_feat = gpt2_oracle.clip_encode_text(built, tspec, torch.tensor(clip_tokens), mode="as_built")
    ora_feat32 = gpt2_oracle.clip_encode_text(built, tspec, torch.tensor(clip_tokens), mode="fp32")
    err = (ora_feat.float() - ref_feat.float()).abs().max().item() / ref_feat.float().abs().max().item()
    print(f"[{name}] oracle encode_text vs reference: rel {err:.3e}")
    assert err < 2e-3, err
    # cached image features (generator.py:26-27): a direction with non-trivial, candidate-dependent cosines
    g = torch.Generator().manual_seed(seed + 4)
    f = ora_feat32
    centre = f.mean(0)
    u = ((torch.randn(P, generator=g)[:, None]) * (f - centre)).sum(0)
    image = (0.8 * centre / centre.norm() + 0.6 * u / u.norm())[None].half()
    sim_ref = torch.cosine_similarity(ref_feat, image)                                              # generator.py:59
    sim32 = gpt2_oracle.text_similarity(ora_feat32, image)
    print(f"[{name}] sim (reference, fp16): {sim_ref.float().numpy()}")
    np.savez_compressed(
        os.path.join(out_dir, f"{name}.npz"),
        pop=P, seed=seed, z=z, init_tokens=np.asarray(init), tokens=ref_tokens, first_logits_gap_min=gap.min(),
        clip_tokens=clip_tokens, image_features=image.numpy(), text_features=ref_feat.float().numpy(),
        sim_fp16=sim_ref.numpy(), sim_oracle_fp32=sim32.numpy().astype(np.float32),
    )
    print(f"[{name}] wrote fixture ({time.time() - t0:.1f}s)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--out", default=os.path.join(REPO, "tests", "golden"))
    args = ap.parse_args()
    for name in FIXTURES:
        if args.only and name != args.only:
            continue
        make_fixture(name, args.out)


if __name__ == "__main__":
    main()
