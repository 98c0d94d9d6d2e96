"""-m gpu: the CUDA path (through the C ABI) against the oracle, the emulation
and the committed reference fixtures.  /root/reference is NOT available on the
GPU box: everything here uses oracle/ + tests/golden only.

Tolerances (north_star: scores within 1e-3 relative of the reference):
  * sim: |ours - oracle_fp32| / |oracle_fp32| <= 1e-3, compared in fp32
    "before the final cast"; against the reference's own fp16-rounded value
    1.5e-3 (its last bit alone is 8e-4 at 0.3).
  * hinge: |ours - reference| <= 8e-4 absolute (values ~0.9: <= 1e-3 relative; achieved errors are recorded in
    profiles/r02_parity_full.json).
  * layer intermediates vs the emulation (identical rounding points):
    <= 3e-3 of the tensor's max (a few fp16 ulps).
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from clip_glass_b200 import weights as W
from tests.fixtures import build_inputs, load_golden

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_diag(config, impl, timeout, flags=0):
    """gpu_diag in a subprocess with a timeout: a trapped or hung kernel fails one test, not the session."""
    cmd = [sys.executable, "-m", "tests.gpu_diag", "--config", config, "--impl", str(impl), "--flags", str(flags)]
    r = subprocess.run(cmd, cwd=REPO, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return {d["name"]: d for d in (json.loads(l) for l in r.stdout.splitlines() if l.startswith("{"))}


def check_diag(lines, layer_tol=3e-3):
    for name, d in lines.items():
        if "rel" in d:
            assert d["nonfinite"] == 0, name
            tol = 2e-3 if name in ("images_vs_oracle", "images_vs_emulation") else layer_tol
            if name in ("w", "styles"):
                tol = 1e-5
            assert d["rel"] <= tol, (name, d)
    assert lines["neg_sim_vs_oracle"]["max_rel"] <= 1e-3, lines["neg_sim_vs_oracle"]
    assert lines["hinge_vs_oracle"]["max_abs"] <= 8e-4, lines["hinge_vs_oracle"]
    assert lines["sim_vs_reference_fixture_fp16"]["max_rel"] <= 1.5e-3
    assert lines["launches"]["count"] > 0


def test_tiny_layerwise_tcgen05_path():
    check_diag(run_diag("tiny", 0, 600))


def test_tiny_layerwise_simt_bringup_path():
    check_diag(run_diag("tiny", 1, 600))


def test_tiny_layerwise_exact_resampling_variant():
    """GLASS_FLAG_EXACT_RESAMPLE: every up/down conv from 16x16 inputs up in its exact polyphase form (2x2-tap conv
    + k_upfir / k_blur_s2d).  With the tiny config's 32/64 channels the default cost model keeps every layer in
    the folded form, so this is the test that exercises the exact kernels."""
    check_diag(run_diag("tiny", 0, 600, flags=2))


def test_tiny_layerwise_fused_projection_variant():
    """GLASS_FLAG_PROJ_FUSION on the tiny config: block 0 runs fromRGB + FIR + projection (two 32-channel chunks),
    block 1 the I8-input variant, block 2 (32 output channels) stays on the separate launches."""
    check_diag(run_diag("tiny", 0, 600, flags=64))


def test_tiny_layerwise_single_pixel_variant():
    """GLASS_FLAG_NO_PAIR_PACK: the 32-channel convs on single pixels (resident-tap MODE 1 with 64-byte rows)
    instead of the default horizontally paired pixels."""
    check_diag(run_diag("tiny", 0, 600, flags=4))


@pytest.fixture(scope="module")
def full_engine():
    from clip_glass_b200.engine import GlassEngine
    inp = build_inputs("full")
    gold = load_golden("full")
    eng = GlassEngine(inp["gan"], inp["clip"], inp["g_sd"], inp["d_sd"], inp["c_sd"], batch_size=inp["batch"],
                      max_population=16)
    eng.set_text_features(torch.from_numpy(gold["text_features"]))
    yield eng, inp, gold
    eng.close()


@pytest.mark.parametrize("flags", [1, 2, 4])
def test_full_size_resampling_variants_agree(flags):
    """ffhq-f at full size with every resampling conv folded (1) / exact (2), or without pixel pairing (4):
    same scores as the fixture."""
    from clip_glass_b200.engine import GlassEngine
    inp = build_inputs("full")
    gold = load_golden("full")
    eng = GlassEngine(inp["gan"], inp["clip"], inp["g_sd"], inp["d_sd"], inp["c_sd"], batch_size=inp["batch"],
                      max_population=4, flags=flags)
    eng.set_text_features(torch.from_numpy(gold["text_features"]))
    neg_sim, hinge = eng.evaluate(inp["x"], noise=inp["noise"])
    eng.close()
    sim32 = gold["sim_oracle_fp32"]
    assert np.abs(-neg_sim - sim32).max() / np.abs(sim32).min() <= 1e-3, (-neg_sim, sim32)
    np.testing.assert_allclose(hinge, gold["F"][:, 1], atol=8e-4)


def test_full_size_against_reference_fixture(full_engine):
    """ffhq-config-f 1024^2 + ViT-B/32, P=4: scores vs the values the UNMODIFIED reference modules
    produced in the build container (tests/golden/full.npz)."""
    eng, inp, gold = full_engine
    neg_sim, hinge = eng.evaluate(inp["x"], noise=inp["noise"])
    sim32 = gold["sim_oracle_fp32"]
    assert np.abs(-neg_sim - sim32).max() / np.abs(sim32).min() <= 1e-3, (-neg_sim, sim32)
    ref16 = gold["sim_fp16"].astype(np.float32)
    assert (np.abs(-neg_sim - ref16) / np.abs(ref16)).max() <= 1.5e-3
    np.testing.assert_allclose(hinge, gold["F"][:, 1], atol=8e-4)
    z = torch.from_numpy(inp["x"]).float().cuda()
    images = eng.generate(z, noise=inp["noise"])
    assert images.shape == (4, 3, 1024, 1024) and float(images.min()) >= 0 and float(images.max()) <= 1
    small = torch.nn.functional.avg_pool2d(images, 16).cpu().numpy()
    np.testing.assert_allclose(small, gold["images_64"], atol=1e-3)
    np.testing.assert_allclose(images.mean(dim=(1, 2, 3)).cpu().numpy(), gold["image_mean"], atol=2e-4)


def test_full_size_properties(full_engine):
    """Size-independent properties at a larger population (P=16 here; bench.py runs P=64):
    minibatch groups are independent, results do not depend on which other groups share the launch,
    group order is equivariant, seeded noise is deterministic and shard-invariant."""
    eng, inp, gold = full_engine
    P, B = 16, inp["batch"]
    x = W.make_latents(P, 512, 77)
    noise = W.make_noise(inp["gan"], P // B, 78)
    f_all, h_all = eng.evaluate(x, noise=noise)
    f_sub, h_sub = eng.evaluate(x[4:12], noise=noise[1:3])
    np.testing.assert_array_equal(f_all[4:12], f_sub)
    np.testing.assert_array_equal(h_all[4:12], h_sub)
    # one minibatch alone: populations below 8 run the per-candidate GEMV kernel instead of the population-tiled one
    # (kernels.cu: vecmat_tile_kernel keeps vecmat_kernel's summation order) and the small-grid tile-width rule picks
    # other column tiles for the 4x4 .. 16x16 layers -- the scores must not notice either
    f_one, h_one = eng.evaluate(x[4:8], noise=noise[1:2])
    np.testing.assert_array_equal(f_all[4:8], f_one)
    np.testing.assert_array_equal(h_all[4:8], h_one)
    perm = np.concatenate([np.arange(8, 12), np.arange(0, 4), np.arange(12, 16), np.arange(4, 8)])
    f_p, h_p = eng.evaluate(x[perm], noise=[noise[2], noise[0], noise[3], noise[1]])
    np.testing.assert_array_equal(f_p, f_all[perm])
    np.testing.assert_array_equal(h_p, h_all[perm])
    # identical latents inside one minibatch share the noise draw -> identical scores; across groups they differ
    xs = np.repeat(x[:1], P, axis=0)
    f_s, h_s = eng.evaluate(xs, noise=noise)
    assert np.all(f_s[:4] == f_s[0]) and not np.all(f_s[4:8] == f_s[0])
    # seeded device noise: deterministic, and a shard evaluated on its own draws the same noise
    a1 = eng.evaluate(x, seed=11)
    a2 = eng.evaluate(x, seed=11)
    a3 = eng.evaluate(x, seed=12)
    np.testing.assert_array_equal(a1[0], a2[0])
    assert not np.array_equal(a1[0], a3[0])
    shard = eng.evaluate(x[8:], seed=11, first_group=2)
    np.testing.assert_array_equal(shard[0], a1[0][8:])
    np.testing.assert_array_equal(shard[1], a1[1][8:])
    # all scores are finite, sims within [-1,1], hinge >= 0
    assert np.isfinite(f_all).all() and np.abs(f_all).max() <= 1.0 and (h_all >= 0).all()


def _full_scores(flags=0, env=None, evals=1):
    """Scores of the full-size fixture from a fresh engine built with `flags` / environment knobs."""
    from clip_glass_b200.engine import GlassEngine
    inp = build_inputs("full")
    gold = load_golden("full")
    old = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        eng = GlassEngine(inp["gan"], inp["clip"], inp["g_sd"], inp["d_sd"], inp["c_sd"], batch_size=inp["batch"],
                          max_population=4, flags=flags)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    eng.set_text_features(torch.from_numpy(gold["text_features"]))
    outs = [eng.evaluate(inp["x"], noise=inp["noise"]) for _ in range(evals)]
    eng.close()
    return outs, gold


def test_tcgen05_attention_against_simt_cross_check():
    """attention_tc.cu (QK^T / PV on the tensor cores, P rounded to fp16 like the reference's fp16 softmax output)
    against the scalar shared-memory kernel (GLASS_FLAG_SIMT_ATTENTION = 16) on the full ViT-B/32: the final cosine
    moves by far less than the 1e-3 budget."""
    (tc,), gold = _full_scores(flags=0)
    (simt,), _ = _full_scores(flags=16)
    np.testing.assert_allclose(tc[0], simt[0], rtol=3e-4)
    np.testing.assert_array_equal(tc[1], simt[1])          # the discriminator does not depend on the CLIP tower
    sim32 = gold["sim_oracle_fp32"]
    assert np.abs(-simt[0] - sim32).max() / np.abs(sim32).min() <= 1e-3


def test_graph_replay_equals_eager_launches():
    """glass_evaluate_host: the first evaluation of a plan launches eagerly, later ones replay a CUDA graph with
    the CLIP tower and the discriminator as parallel branches -- bit-identical scores, and identical to an engine
    that never builds a graph (GLASS_FLAG_NO_GRAPH = 32)."""
    outs, _ = _full_scores(flags=0, evals=4)
    for o in outs[1:]:
        np.testing.assert_array_equal(o[0], outs[0][0])
        np.testing.assert_array_equal(o[1], outs[0][1])
    (eager,), _ = _full_scores(flags=32)
    np.testing.assert_array_equal(eager[0], outs[0][0])
    np.testing.assert_array_equal(eager[1], outs[0][1])


def test_streamed_tap_i8_downconv_against_nhwc_form():
    """conv_tc MODE 6 (D 512^2 folded down-conv on the I8 space-to-depth tensor, taps streamed through a ring)
    against the same layer on the NHWC tensor with MODE 0 (GLASS_FLAG_C1_NHWC = 128): same products, different
    accumulation order only."""
    (m6,), gold = _full_scores()
    (m0,), _ = _full_scores(flags=128)
    np.testing.assert_array_equal(m6[0], m0[0])            # G and CLIP are untouched
    np.testing.assert_allclose(m6[1], m0[1], atol=5e-4)
    np.testing.assert_allclose(m6[1], gold["F"][:, 1], atol=8e-4)


def test_fused_fir_downconv_against_folded_form():
    """downconv_tc.cu (default for the 32/64-channel D blocks: exact 3x3 stride-2 conv, FIR applied inside the kernel
    in packed-half2 arithmetic) against the FIR-folded space-to-depth form (GLASS_FLAG_NO_FUSED_DOWN = 512, 4x the
    MACs, fp32 accumulation of the folded taps): G and CLIP are untouched, the hinge agrees to fp16 rounding of the
    blurred operand, and both stay within the bound against the reference fixture."""
    (fused,), gold = _full_scores(flags=0)
    (folded,), _ = _full_scores(flags=512)
    np.testing.assert_array_equal(fused[0], folded[0])
    np.testing.assert_allclose(fused[1], folded[1], atol=5e-4)
    np.testing.assert_allclose(fused[1], gold["F"][:, 1], atol=8e-4)
    np.testing.assert_allclose(folded[1], gold["F"][:, 1], atol=8e-4)


def test_projection_accumulator_against_separate_gemm():
    """The 32 -> 64 discriminator block computes its projection path (1x1 conv of the FIR-downsampled block input,
    stylegan2/modules.py:1352-1372) as a second TMEM accumulator of the fused-FIR down-conv kernel, in fp32, instead of
    a separate GEMM whose fp16 output is read back as the residual (GLASS_FLAG_NO_PROJ_ACC = 4096): G and CLIP are
    untouched, the hinge agrees to the fp16 rounding of that residual tensor, both within the bound of the fixture."""
    (acc,), gold = _full_scores(flags=0)
    (sep,), _ = _full_scores(flags=4096)
    np.testing.assert_array_equal(acc[0], sep[0])
    np.testing.assert_allclose(acc[1], sep[1], atol=5e-4)
    np.testing.assert_allclose(acc[1], gold["F"][:, 1], atol=8e-4)
    np.testing.assert_allclose(sep[1], gold["F"][:, 1], atol=8e-4)


def test_image_finished_in_last_conv_epilogue_against_rgb_combine():
    """The last generator conv finishes the image in its epilogue (skip sum + x2 upsample + toRGB bias + biggan_norm)
    instead of writing a toRGB slab for k_rgb_combine (GLASS_FLAG_NO_IMAGE_FUSION = 1024): the same fp32 arithmetic in
    the same order, so images agree to the last bits and the scores to far below the parity bound."""
    from clip_glass_b200.engine import GlassEngine
    inp, gold = build_inputs("full"), load_golden("full")
    z = torch.from_numpy(inp["x"]).float().cuda()
    outs = []
    for flags in (0, 1024):
        eng = GlassEngine(inp["gan"], inp["clip"], inp["g_sd"], inp["d_sd"], inp["c_sd"], batch_size=inp["batch"],
                          max_population=4, flags=flags)
        eng.set_text_features(torch.from_numpy(gold["text_features"]))
        img = eng.generate(z, noise=inp["noise"])
        outs.append((img.clone(), eng.evaluate(inp["x"], noise=inp["noise"])))
        eng.close()
    (img_f, (s_f, h_f)), (img_c, (s_c, h_c)) = outs
    assert float((img_f - img_c).abs().max()) <= 2e-6
    np.testing.assert_allclose(s_f, s_c, rtol=2e-5)
    np.testing.assert_allclose(h_f, h_c, atol=2e-5)
    small = torch.nn.functional.avg_pool2d(img_f, 16).cpu().numpy()
    np.testing.assert_allclose(small, gold["images_64"], atol=1e-3)


def test_fused_projection_kernel_against_separate_launches():
    """fir_proj_tc.cu (GLASS_FLAG_PROJ_FUSION = 64: stride-2 FIR + 1x1 projection GEMM of the D blocks in one kernel,
    opt-in) against the default k_fir_down + conv_tc route: the projection accumulates in fp32 from the same fp16
    operands either way, so the hinge agrees to fp16 rounding of the residual tensor."""
    (fused,), gold = _full_scores(flags=64)
    (ref,), _ = _full_scores(flags=0)
    np.testing.assert_array_equal(fused[0], ref[0])
    np.testing.assert_allclose(fused[1], ref[1], atol=5e-4)
    np.testing.assert_allclose(fused[1], gold["F"][:, 1], atol=8e-4)


# ---- round 2: parity at more seeds, at the benchmarked plan, under fp16-range stress ------------------------------
# Bounds (north_star: fitness within 1e-3 of the reference).  sim: relative, against the fp32 oracle value.  hinge =
# relu(1 - d) with |d| <= ~0.2 on these fixtures, so a bound relative to the hinge (~0.9) and a bound relative to the
# logit differ by 5-20x: the test bounds the ABSOLUTE error by HINGE_ATOL, i.e. <= 1e-3 relative to the hinge value,
# and profiles/r02_parity_full.json records the achieved error next to the largest logit of each fixture.
SIM_RTOL = 1e-3
HINGE_ATOL = 8e-4


@pytest.mark.parametrize("name", ["full", "full_b", "full_c"])
def test_full_size_fixtures_multi_seed(name):
    """Three weight / latent / noise seeds at ffhq-config-f 1024^2 + ViT-B/32; full_c has two noise groups and both
    MinibatchStd pairings (P = 8)."""
    from tests.parity_report import fixture_record
    rec = fixture_record(name, with_images=True, range_check=True)
    assert rec["finite"] and rec["range"]["nonfinite"] == 0 and rec["range"]["saturated"] == 0, rec
    assert rec["sim_rel_vs_oracle_fp32"] <= SIM_RTOL, rec
    assert rec["sim_rel_vs_reference_fp16"] <= 1.5e-3, rec
    assert rec["hinge_abs"] <= HINGE_ATOL, rec
    assert rec["image64_abs"] <= 1e-3 and rec["image_mean_abs"] <= 2e-4, rec


def test_full_size_single_objective_config():
    """StyleGAN2_ffhq_nod (no discriminator, F = -sim) at full size against the fixture's F_nod."""
    from tests.parity_report import fixture_record
    rec = fixture_record("full_b", use_d=False)
    assert rec["finite"] and rec["sim_rel_vs_oracle_fp32"] <= SIM_RTOL and rec["sim_rel_vs_reference_fp16"] <= 1.5e-3, rec


def test_benchmarked_plan_p64_embeds_golden_groups():
    """max_population = P = 64, 16 noise groups: the plan bench.py times (tile decode, tiles_n and the pow2 paths
    depend on P).  Golden groups sit at slots 0, 7 and 15 among filler candidates; their scores equal the fixture,
    the CUDA-graph replay equals the eager first evaluation, and the same candidates in a P = 8 plan are
    bit-identical."""
    from tests.parity_report import embedded_p64_record
    rec = embedded_p64_record()
    assert rec["all_finite"] and rec["graph_equals_eager"] and rec["equals_small_plan_bitwise"], rec
    assert rec["sim_rel_vs_oracle_fp32"] <= SIM_RTOL and rec["hinge_abs"] <= HINGE_ATOL, rec


def test_fp16_range_stress_tiny_layerwise():
    """Style biases x30 / x20000 and conv weights x10 on a few channels (weights.make_generator_weights(stress=True)):
    activation x style exceeds 65504 unless the styles are normalised (k_style_norm).  Layer by layer against the
    emulation and the oracle, on the tensor-core path."""
    check_diag(run_diag("tiny_stress", 0, 600))


def test_fp16_range_stress_full_size():
    """The same stress at ffhq-config-f size: scores within the north-star bound, zero non-finite and zero saturated
    fp16 activations over every G and D layer (glass_set_range_check)."""
    from tests.parity_report import fixture_record
    rec = fixture_record("full_stress", range_check=True)
    assert rec["finite"] and rec["range"]["nonfinite"] == 0 and rec["range"]["saturated"] == 0, rec
    assert rec["range"]["max_abs"] < 6.0e4, rec
    assert rec["sim_rel_vs_oracle_fp32"] <= SIM_RTOL and rec["hinge_abs"] <= HINGE_ATOL, rec


def test_device_noise_generator_is_standard_normal():
    """noise_kernel (Philox4x32-10 + Box-Muller) replaces the reference's ``normal_()`` draw per minibatch forward
    (stylegan2/modules.py:426-452): mean, variance, skewness, kurtosis and tails of ~700k draws, a KS test against
    N(0,1), and independence of consecutive groups."""
    from scipy import stats
    from clip_glass_b200.engine import GlassEngine
    gan, clip = W.TINY_GAN, W.TINY_CLIP
    P, B = 256, 4
    eng = GlassEngine(gan, clip, W.make_generator_weights(gan, 1), None, W.make_clip_visual_weights(clip, 2),
                      batch_size=B, max_population=P)
    eng.set_debug(capture=True)
    z = torch.from_numpy(W.make_latents(P, gan.latent_size, 3)).float().cuda()
    eng.generate(z, seed=12345)
    nz = eng.debug_read("noise").astype(np.float64)
    eng.close()
    n = nz.size
    assert n == (P // B) * sum(r * r for r in gan.noise_shapes())
    assert np.isfinite(nz).all()
    assert abs(nz.mean()) < 5.0 / np.sqrt(n)
    assert abs(nz.var() - 1.0) < 5.0 * np.sqrt(2.0 / n)
    assert abs(stats.skew(nz)) < 5.0 * np.sqrt(6.0 / n)
    assert abs(stats.kurtosis(nz)) < 5.0 * np.sqrt(24.0 / n)
    for k, p in ((2.0, 0.04550026), (3.0, 0.00269980), (4.0, 6.334e-5)):
        frac = (np.abs(nz) > k).mean()
        assert abs(frac - p) < 5.0 * np.sqrt(p / n) + 1e-7, (k, frac, p)
    assert stats.kstest(nz[:200000], "norm").pvalue > 1e-4
    g = nz.reshape(P // B, -1)
    c = np.corrcoef(g[:-1].reshape(-1)[:500000], g[1:].reshape(-1)[:500000])[0, 1]
    assert abs(c) < 5.0 / np.sqrt(500000)


# ---- round 2: the callers either side of the path ---------------------------------------------
def test_image_output_path_reuses_scored_images(tmp_path):
    """§8(f)-4.  run.py:29-51 save_callback -> generator.generate -> generator.save.  After a fused _evaluate the
    engine still holds the images it scored: generate() returns exactly those for the rows it finds (here: equal,
    bit for bit, to a render with the same explicit noise), renders only the rest, and save() assembles the grid +
    uint8 conversion on the device, byte-identical to torchvision's make_grid + save_image arithmetic."""
    import torchvision
    from PIL import Image
    from clip_glass_b200.config import make_namespace
    from clip_glass_b200.problem import GenerationProblem
    gold, inp = load_golden("tiny"), build_inputs("tiny")
    ns = make_namespace("StyleGAN2_ffhq_d", device="cuda:0", pop_size=8, batch_size=4, synthetic_seed=100,
                        gan_spec=W.TINY_GAN, clip_spec=W.TINY_CLIP, text_features=torch.from_numpy(gold["text_features"]))
    prob = GenerationProblem(ns)
    gen = prob.generator
    out = {}
    prob._evaluate(inp["x"], out, noise=inp["noise"])
    z = torch.from_numpy(inp["x"]).float().cuda()
    rendered = gen.engine.generate(z, noise=inp["noise"])                 # the same candidates, same noise
    order = [5, 2, 7, 0, 1, 3]
    ls = ns.latent(ns)
    ls.set_from_population(inp["x"][order])
    got = gen.generate(ls, minibatch=2)
    assert gen.reuse_stats == dict(reused=6, rendered=0)
    assert torch.equal(got, rendered[order])
    # two known rows + two new candidates: only the new ones are rendered
    mixed = np.concatenate([inp["x"][[4, 6]], W.make_latents(2, 512, 9)])
    ls.set_from_population(mixed)
    got2 = gen.generate(ls, minibatch=4)
    assert gen.reuse_stats == dict(reused=8, rendered=2) and got2.shape == (4, 3, 64, 64)
    assert torch.equal(got2[:2], rendered[[4, 6]]) and float(got2[2:].min()) >= 0 and float(got2[2:].max()) <= 1
    # grid + uint8 conversion: bytes equal torchvision's
    for n, nrow in ((6, 8), (8, 3), (1, 8)):
        imgs = rendered[:n].contiguous()
        ours = gen.engine.image_grid_u8(imgs, nrow=nrow, padding=2 if n > 1 else 0)
        grid = torchvision.utils.make_grid(imgs.cpu(), nrow=nrow) if n > 1 else imgs[0].cpu()
        ref = grid.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(torch.uint8).numpy()
        np.testing.assert_array_equal(ours, ref)
    path = tmp_path / "grid.jpg"
    gen.save(rendered, str(path))
    with Image.open(path) as im:
        assert im.size == (8 * 66 + 2, 66 + 2)
    with pytest.raises(Exception):
        gen.engine.last_images([99])


def test_biggan_latent_arithmetic_on_device():
    """§8(f)-3, latent.py:16-24: z = clip(x[:, :128], -2, 2), class vector = softmax over the 1000 'bool' genes, from
    the float64 mixed population as pymoo hands it over (glass_biggan_latent) against the reference's torch ops."""
    from clip_glass_b200.config import make_namespace
    ns = make_namespace("DeepMindBigGAN512", device="cuda:0")
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.normal(0, 1.5, size=(32, 128)), (rng.random((32, 1000)) < 0.005).astype(float)], axis=1)
    x[3, 128:] = 0.0                                                       # no class gene set: uniform softmax
    ls = ns.latent(ns)
    ls.set_from_population(x)
    z, cl = ls()
    zr = torch.clip(torch.tensor(x[:, :128].astype(float)).float(), -2, 2)
    cr = torch.softmax(torch.tensor(x[:, 128:].astype(float)).float(), dim=1)
    assert torch.equal(z.cpu(), zr)
    np.testing.assert_allclose(cl.cpu().numpy(), cr.numpy(), rtol=2e-6, atol=0)
    np.testing.assert_allclose(cl.sum(1).cpu().numpy(), 1.0, rtol=1e-5)


@pytest.mark.parametrize("config,algorithm_files", [("StyleGAN2_ffhq_nod", 5), ("StyleGAN2_ffhq_d", 3)])
def test_run_driver_end_to_end(tmp_path, config, algorithm_files):
    """BASELINE config 1 (StyleGAN2_ffhq_nod, pop 8, 5 generations) through the driver mirror of run.py:15-125 —
    on the GPU path, reduced network shapes — and the NSGA-II config: sampling -> _evaluate -> mating -> survival ->
    save_callback (images reused from the last evaluation) -> result files."""
    import pickle
    from clip_glass_b200 import run as driver
    gens = algorithm_files
    res = driver.main(["--device", "cuda:0", "--config", config, "--generations", str(gens), "--save-each", "2",
                       "--tmp-folder", str(tmp_path), "--pop-size", "8", "--batch-size", "4", "--synthetic-seed", "100",
                       "--seed", "3"], config_overrides=dict(gan_spec=W.TINY_GAN, clip_spec=W.TINY_CLIP))
    names = set(os.listdir(tmp_path))
    assert {"genetic-it-2.jpg", "genetic-it-final.jpg", "output.jpg", "genetic_result", "ls_result"} <= names, names
    with open(tmp_path / "genetic_result", "rb") as f:
        saved = pickle.load(f)
    F = np.atleast_2d(np.asarray(saved["F"], dtype=float))
    assert np.isfinite(F).all() and len(res.pop) == 8
    assert "z" in torch.load(tmp_path / "ls_result")


def test_population_must_be_multiple_of_batch(full_engine):
    eng, inp, gold = full_engine
    with pytest.raises(AssertionError):            # models.py:112,124
        eng.evaluate(W.make_latents(6, 512, 1), seed=1)
    with pytest.raises(AssertionError):
        eng.evaluate(W.make_latents(32, 512, 1), seed=1)      # > max_population


def test_plugin_surface_fused_equals_facade_calls():
    """GenerationProblem._evaluate: the one-call fused route and the reference's three façade calls
    (generate -> clip_similarity -> discriminate) give the same F; shapes/dtypes as pymoo expects."""
    from clip_glass_b200.config import make_namespace
    from clip_glass_b200.problem import GenerationProblem
    gold = load_golden("tiny")
    inp = build_inputs("tiny")
    ns = make_namespace("StyleGAN2_ffhq_d", device="cuda:0", pop_size=8, batch_size=4, synthetic_seed=100,
                        gan_spec=W.TINY_GAN, clip_spec=W.TINY_CLIP,
                        text_features=torch.from_numpy(gold["text_features"]))
    prob = GenerationProblem(ns)
    out_f, out_u = {}, {}
    prob._evaluate(inp["x"], out_f, noise=inp["noise"])
    ns.fused = False
    ls = ns.latent(ns)
    ls.set_from_population(inp["x"])
    images = prob.generator.generate(ls, minibatch=4, noise=inp["noise"])
    sim = prob.generator.clip_similarity(images)
    dis = prob.generator.discriminate(images, minibatch=4)
    assert images.shape == (8, 3, 64, 64) and sim.shape == (8,) and dis.shape == (8, 1)
    F_u = np.column_stack((-sim.cpu().numpy(), torch.relu(1 - dis).squeeze(1).cpu().numpy()))
    assert out_f["F"].shape == (8, 2) and out_f["G"].shape == (8,) and not out_f["G"].any()
    np.testing.assert_array_equal(out_f["F"], F_u)
    np.testing.assert_allclose(out_f["F"][:, 0], gold["F"][:, 0], rtol=1.5e-3)
    np.testing.assert_allclose(out_f["F"][:, 1], gold["F"][:, 1], atol=8e-4)
    assert prob.generator.has_discriminator()
    # minibatch=None == one noise draw for the whole call (generator.py:29-31 / models.py:109-110), any P
    ls.set_from_population(inp["x"][:3])
    assert prob.generator.generate(ls).shape == (3, 3, 64, 64)
    # single-objective config: F is [P]
    ns1 = make_namespace("StyleGAN2_ffhq_nod", device="cuda:0", pop_size=8, batch_size=4, synthetic_seed=100,
                         gan_spec=W.TINY_GAN, clip_spec=W.TINY_CLIP,
                         text_features=torch.from_numpy(gold["text_features"]))
    p1 = GenerationProblem(ns1)
    o1 = {}
    p1._evaluate(inp["x"], o1, noise=inp["noise"])
    assert o1["F"].shape == (8,)
    np.testing.assert_allclose(o1["F"], gold["F_nod"], rtol=1.5e-3)


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()
