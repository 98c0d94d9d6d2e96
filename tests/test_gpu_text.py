"""-m gpu: the img2txt path (BASELINE config 5) through the C ABI (glass_text_*) against the
fixtures that oracle/make_golden_gpt2.py wrote from the UNMODIFIED reference modules (gpt2.model + gpt2.sample,
clip.model.CLIP.encode_text) and against the oracle restatement.

Bars: GPT-2 token output is integer work -> bit-exact (the GEMMs run as split-fp16 tensor-core products with fp32
accumulation, see text_engine.cu).  CLIP text cosine: <= 1e-3 relative against the fp32-arithmetic oracle value
("before the final cast"), <= 3e-3 against the reference's own fp16 value (its fp16 MultiheadAttention rounds the
scores and probabilities; one fp16 ulp of a cosine near 0.8 is already 6e-4).
"""
import os

import numpy as np
import pytest
import torch

from clip_glass_b200 import text_weights as TW

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# must match oracle/make_golden_gpt2.py FIXTURES
CONFIGS = {
    "gpt2_tiny": dict(gpt2=TW.TINY_GPT2, text=TW.TINY_CLIP_TEXT, pop=8, seed=700),
    "gpt2_full": dict(gpt2=TW.GPT2_SMALL, text=TW.CLIP_TEXT_B32, pop=8, seed=800),
}


def _engine(name, max_population=None, flags=0):
    from clip_glass_b200.text_engine import TextEngine
    cfg = CONFIGS[name]
    gold = dict(np.load(os.path.join(REPO, "tests", "golden", f"{name}.npz")))
    g_sd = TW.make_gpt2_weights(cfg["gpt2"], cfg["seed"])
    t_sd = TW.make_clip_text_weights(cfg["text"], cfg["seed"] + 1)
    eng = TextEngine(cfg["gpt2"], g_sd, cfg["text"], t_sd, init_tokens=gold["init_tokens"].tolist(), dim_z=20,
                     max_tokens_len=30, max_population=max_population or cfg["pop"], flags=flags)
    eng.set_image_features(torch.from_numpy(gold["image_features"]))
    return eng, gold, g_sd, t_sd, cfg


@pytest.mark.parametrize("name", ["gpt2_tiny", "gpt2_full"])
def test_gpt2_greedy_decode_tokens_are_bit_exact(name):
    eng, gold, g_sd, t_sd, cfg = _engine(name)
    tokens = eng.generate_tokens(gold["z"])
    assert tokens.shape == gold["tokens"].shape and tokens.dtype == np.int64
    np.testing.assert_array_equal(tokens[:, :23], gold["tokens"][:, :23])          # cat(z, init tokens)
    first_bad = np.argmax(tokens != gold["tokens"], axis=1)
    assert np.array_equal(tokens, gold["tokens"]), (first_bad, tokens[:, 23:27], gold["tokens"][:, 23:27])
    assert np.array_equal(eng.generate_tokens(gold["z"]), tokens)                  # deterministic
    # candidates are independent: a sub-population gives the same rows
    np.testing.assert_array_equal(eng.generate_tokens(gold["z"][2:5]), tokens[2:5])
    eng.close()


def test_gpt2_decode_split_k_equals_single_pass_gemms():
    """The decode-step GEMMs run split-K with the reduction fused into the consuming kernels; the cross-check variant
    (GLASS_TEXT_FLAG_NO_SPLIT_K: one CTA per n-tile over the whole K, epilogue-fused bias / GELU / residual) sums the
    same products in another order.  Both reproduce the reference's tokens."""
    for flags in (0, 1):
        eng, gold, g_sd, t_sd, cfg = _engine("gpt2_full", flags=flags)
        np.testing.assert_array_equal(eng.generate_tokens(gold["z"]), gold["tokens"])
        eng.close()


def test_gpt2_decode_at_the_benchmarked_population():
    """P = 64 (BASELINE config 5): the fixture's 8 candidates among 56 others decode to the same tokens, and fresh
    candidates match the oracle restatement run on the CPU for a few of them."""
    from oracle import gpt2_oracle
    eng, gold, g_sd, t_sd, cfg = _engine("gpt2_full", max_population=64)
    z = TW.make_token_latents(64, 20, cfg["gpt2"].vocab, 4321)
    z[8:16] = gold["z"]
    tokens = eng.generate_tokens(z)
    np.testing.assert_array_equal(tokens[8:16], gold["tokens"])
    ref = gpt2_oracle.gpt2_generate_tokens(g_sd, cfg["gpt2"], z[40:44], gold["init_tokens"].tolist(), 30)
    np.testing.assert_array_equal(tokens[40:44], ref)
    eng.close()


@pytest.mark.parametrize("name", ["gpt2_tiny", "gpt2_full"])
def test_clip_text_similarity_against_reference_fixture(name):
    eng, gold, g_sd, t_sd, cfg = _engine(name)
    sim, feats = eng.text_similarity(gold["clip_tokens"], return_features=True)
    s32, s16 = gold["sim_oracle_fp32"], gold["sim_fp16"].astype(np.float32)
    assert np.isfinite(sim).all()
    assert (np.abs(sim - s32) / np.abs(s32)).max() <= 1e-3, (sim, s32)
    assert (np.abs(sim - s16) / np.abs(s16)).max() <= 3e-3, (sim, s16)
    scale = np.abs(gold["text_features"]).max()
    assert np.abs(feats - gold["text_features"]).max() / scale <= 5e-3
    # rows are independent and the EOT position (arg-max token) is what selects the feature row
    np.testing.assert_array_equal(eng.text_similarity(gold["clip_tokens"][3:6]), sim[3:6])
    eng.close()


def test_text_engine_argument_errors():
    eng, gold, g_sd, t_sd, cfg = _engine("gpt2_tiny")
    bad = gold["z"].copy()
    bad[0, 0] = cfg["gpt2"].vocab                       # the reference's embedding lookup raises IndexError
    with pytest.raises(AssertionError):
        eng.generate_tokens(bad)
    with pytest.raises(AssertionError):
        eng.generate_tokens(np.repeat(gold["z"], 2, axis=0))     # > max_population
    ct = gold["clip_tokens"].copy()
    ct[0, 1] = cfg["text"].vocab
    with pytest.raises(AssertionError):
        eng.text_similarity(ct)
    eng.close()


def test_plugin_surface_gpt2_config():
    """GenerationProblem._evaluate for config GPT2 (problem.py:14-29 on the img2txt task): x holds integer genes as
    float64 (pymoo), F = -sim [P], G = zeros; the token-level route (no vocabulary files on this box) equals the two
    engine calls made by hand; Generator.save writes the texts file (generator.py:69-71)."""
    from clip_glass_b200.config import make_namespace
    from clip_glass_b200.models import standin_clip_tokens
    from clip_glass_b200.problem import GenerationProblem
    cfg = CONFIGS["gpt2_tiny"]
    gold = dict(np.load(os.path.join(REPO, "tests", "golden", "gpt2_tiny.npz")))
    ns = make_namespace("GPT2", device="cuda:0", pop_size=8, batch_size=8, synthetic_seed=cfg["seed"],
                        gpt2_spec=cfg["gpt2"], clip_text_spec=cfg["text"], clip_token_map="standin",
                        image_features=torch.from_numpy(gold["image_features"]))
    prob = GenerationProblem(ns)
    assert not prob.generator.has_discriminator()
    out = {}
    prob._evaluate(gold["z"].astype(np.float64), out)
    assert out["F"].shape == (8,) and out["G"].shape == (8,) and not out["G"].any()
    toks = prob.generator.model.generate_tokens(gold["z"])
    np.testing.assert_array_equal(toks, gold["tokens"])               # same seeds as the fixture: same weights
    gen = prob.generator.model.parse_out_tokens(toks)
    assert gen[1] == []                                               # EOT gene inside the latent part (models.py:35-36)
    sim = prob.generator.clip_similarity(standin_clip_tokens(gen, cfg["text"]))
    np.testing.assert_array_equal(out["F"], -sim.numpy())
    ls = ns.latent(ns)
    ls.set_from_population(gold["z"].astype(np.float64))
    assert ls()[0].dtype == torch.int64
    ns.return_tokens = True
    assert prob.generator.generate(ls) == gen
    with pytest.raises(Exception):                                    # text output needs the vocabulary files
        ns.return_tokens = False
        prob.generator.generate(ls)
