"""The oracle (oracle/*.py, a CPU restatement) against the golden fixtures
that oracle/make_golden.py produced by running the UNMODIFIED reference
modules (stylegan2.models.Generator/Discriminator, clip.model.CLIP) in the
build container.  This is what pins the oracle; the CUDA path is then pinned
to the oracle by the -m gpu tests."""
import numpy as np
import pytest
import torch

from clip_glass_b200 import weights as W
from oracle import evaluate_oracle, stylegan2_oracle
from tests.fixtures import build_inputs, load_golden


def _run(name, use_d=True, clip_mode="as_built"):
    inp = build_inputs(name)
    gold = load_golden(name)
    text = torch.from_numpy(gold["text_features"])
    out = evaluate_oracle.evaluate(
        inp["x"], inp["g_sd"], inp["d_sd"], W.clip_as_built(inp["c_sd"]), text,
        inp["gan"], inp["clip"], inp["batch"], use_d, noise=inp["noise"],
        clip_mode=clip_mode, return_images=True)
    return inp, gold, out


def test_tiny_matches_reference_fixture():
    inp, gold, out = _run("tiny")
    # images: fp32 arithmetic re-association only
    small = torch.nn.functional.avg_pool2d(out["images"], max(1, inp["gan"].resolution // 64))
    np.testing.assert_allclose(small.numpy(), gold["images_64"], atol=2e-5)
    np.testing.assert_allclose(out["images"].mean(dim=(1, 2, 3)).numpy(), gold["image_mean"], atol=1e-5)
    # F: column 0 is -sim (fp16 in the reference: one ulp at 0.35 is 2.4e-4), column 1 hinge (fp32)
    assert out["F"].shape == (inp["pop"], 2)
    np.testing.assert_allclose(out["F"][:, 0], gold["F"][:, 0], atol=5e-4)
    np.testing.assert_allclose(out["F"][:, 1], gold["F"][:, 1], rtol=1e-4, atol=1e-5)
    np.testing.assert_array_equal(out["G"], np.zeros(inp["pop"]))


def test_tiny_single_objective_shape():
    inp, gold, out = _run("tiny", use_d=False)
    assert out["F"].shape == (inp["pop"],)
    np.testing.assert_allclose(out["F"], gold["F_nod"], atol=5e-4)


def test_tiny_fp32_clip_mode_close_to_as_built():
    inp, gold, out = _run("tiny", clip_mode="fp32")
    # fp32 arithmetic over the same fp16 weights: within a few fp16 ulps of the as-built score
    np.testing.assert_allclose(-out["F"][:, 0], gold["sim_fp16"].astype(np.float32), rtol=3e-3)
    np.testing.assert_allclose(-out["F"][:, 0], gold["sim_oracle_fp32"], rtol=1e-5, atol=1e-6)


def test_population_must_divide_batch():
    inp = build_inputs("tiny")
    with pytest.raises(AssertionError):     # models.py:112
        evaluate_oracle.generate(torch.zeros(6, 512), inp["g_sd"], inp["gan"], 4, None)


def test_noise_is_shared_within_minibatch_only():
    """modules.py:426-452: noise is [1,1,H,W] per forward => identical latents in
    one minibatch give identical images; different minibatches differ."""
    inp = build_inputs("tiny")
    z = torch.zeros(8, 512)
    z[:] = torch.from_numpy(inp["x"][0]).float()
    imgs = evaluate_oracle.generate(z, inp["g_sd"], inp["gan"], 4, inp["noise"])
    assert torch.equal(imgs[0], imgs[3])
    assert not torch.equal(imgs[0], imgs[4])


def test_mbstd_couples_candidates_and_centres_features():
    """modules.py:726-746 incl. the in-place aliasing quirk documented in the oracle."""
    x = torch.randn(8, 16, 4, 4)
    y = stylegan2_oracle.minibatch_std(x, 4)
    assert y.shape == (8, 17, 4, 4)
    grp = x.reshape(4, 2, 16, 4, 4)
    np.testing.assert_allclose(y[:, :16].reshape(4, 2, 16, 4, 4).numpy(),
                               (grp - grp.mean(0, keepdim=True)).numpy(), atol=1e-6)
    # members j and j+2 (B/G = 2) share the std feature
    assert torch.equal(y[0, 16], y[2, 16]) and not torch.equal(y[0, 16], y[1, 16])


@pytest.mark.slow
def test_full_matches_reference_fixture():
    inp, gold, out = _run("full")
    small = torch.nn.functional.avg_pool2d(out["images"], inp["gan"].resolution // 64)
    np.testing.assert_allclose(small.numpy(), gold["images_64"], atol=2e-5)
    np.testing.assert_allclose(out["F"][:, 0], gold["F"][:, 0], atol=5e-4)
    np.testing.assert_allclose(out["F"][:, 1], gold["F"][:, 1], rtol=1e-4, atol=1e-5)
