"""CPU-side tests (-m "not gpu"): packing algebra via the emulation, C-ABI
library surface, host mirror modules, multi-rank sharding over gloo."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from clip_glass_b200 import _lib, config as cfgmod, dist, operators, packing, weights as W
from oracle import evaluate_oracle
from tests import emulate as E
from tests.fixtures import build_inputs, load_golden

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tiny():
    inp = build_inputs("tiny")
    gold = load_golden("tiny")
    pk = {}
    pk.update(packing.pack_generator(inp["g_sd"], inp["gan"]))
    pk.update(packing.pack_discriminator(inp["d_sd"], inp["gan"]))
    pk.update(packing.pack_clip_visual(inp["c_sd"], inp["clip"]))
    text = torch.from_numpy(gold["text_features"])
    ref = evaluate_oracle.evaluate(inp["x"], inp["g_sd"], inp["d_sd"], W.clip_as_built(inp["c_sd"]), text,
                                   inp["gan"], inp["clip"], inp["batch"], True, noise=inp["noise"],
                                   clip_mode="fp32", return_images=True)
    return inp, gold, pk, text, ref


def test_packed_algorithm_matches_oracle(tiny):
    """Phase-folded up-conv, pre-scaled inputs + demod on the accumulator, fused toRGB,
    polyphase skip upsample, space-to-depth down-conv and the mbstd quirk reproduce the
    oracle; the only difference left is fp16 storage."""
    inp, gold, pk, text, ref = tiny
    z = evaluate_oracle.latents_from_population(inp["x"])
    img = E.emu_generator(pk, inp["gan"], z, inp["noise"], inp["batch"], fp16=True)
    assert (img - ref["images"]).abs().max() < 2e-3
    logits = E.emu_discriminator(pk, inp["gan"], img, inp["batch"])
    np.testing.assert_allclose(torch.relu(1 - logits).numpy(), ref["F"][:, 1], atol=2e-3)
    feats, sim = E.emu_clip(pk, inp["clip"], img, text)
    np.testing.assert_allclose(sim.numpy(), -ref["F"][:, 0], rtol=1e-3)


def test_fold_upconv_is_exact_in_fp32():
    """conv_transpose2d(stride 2) + FIR(pad 1) == 3x3 conv with 4*Cout phases + depth-to-space."""
    from oracle import stylegan2_oracle as so
    torch.manual_seed(0)
    x = torch.randn(2, 8, 5, 5)
    w = torch.randn(6, 8, 3, 3)
    ref = so.fir(torch.nn.functional.conv_transpose2d(x, w.transpose(0, 1), stride=2), so.fir_kernel(1.0, 2), 1, 1)
    got = E.depth_to_space(E.conv_taps(x.permute(0, 2, 3, 1), packing.fold_upconv(w)), 6).permute(0, 3, 1, 2)
    np.testing.assert_allclose(got.numpy(), ref.numpy(), atol=1e-4)


def test_fold_downconv_is_exact_in_fp32():
    """FIR(pad 2) + 3x3 stride-2 conv == 3x3 conv over the space-to-depth input."""
    from oracle import stylegan2_oracle as so
    torch.manual_seed(1)
    x = torch.randn(2, 4, 8, 8)
    w = torch.randn(6, 4, 3, 3)
    ref = torch.nn.functional.conv2d(so.fir(x, so.fir_kernel(1.0, 1), 2, 2), w, stride=2)
    got = E.conv_taps(E.space_to_depth(x.permute(0, 2, 3, 1)), packing.fold_downconv(w)).permute(0, 3, 1, 2)
    np.testing.assert_allclose(got.numpy(), ref.numpy(), atol=1e-4)


def test_pair_packed_conv_is_the_same_conv():
    """[H][W][C] read as [H][W/2][2C] with packing.pair_pack weights == the plain 3x3 conv."""
    torch.manual_seed(2)
    x = torch.randn(2, 6, 8, 4)                       # NHWC
    w = torch.randn(5, 4, 3, 3)
    ref = E.conv_taps(x, packing.taps_plain(w))       # [2,6,8,5]
    got = E.conv_taps(x.reshape(2, 6, 4, 8), packing.pair_pack(w)).reshape(2, 6, 8, 5)
    np.testing.assert_allclose(got.numpy(), ref.numpy(), atol=1e-5)


def test_exact_polyphase_forms_match_reference_ops():
    from oracle import stylegan2_oracle as so
    torch.manual_seed(3)
    x = torch.randn(2, 8, 16, 16)
    w = torch.randn(6, 8, 3, 3)
    ref = so.fir(torch.nn.functional.conv_transpose2d(x, w.transpose(0, 1), stride=2), so.fir_kernel(1.0, 2), 1, 1)
    u = E.depth_to_space(E.conv_taps_table(x.permute(0, 2, 3, 1), packing.exact_upconv(w), packing.UP_EXACT_TAPS,
                                           (17, 17)), 6)
    got = E.fir_same(u, packing.F1_UP, (32, 32), 1).permute(0, 3, 1, 2)
    np.testing.assert_allclose(got.numpy(), ref.numpy(), atol=1e-4)
    a = torch.randn(2, 4, 32, 32)
    wd = torch.randn(6, 4, 3, 3)
    refd = torch.nn.functional.conv2d(so.fir(a, so.fir_kernel(1.0, 1), 2, 2), wd, stride=2)
    ub = E.fir_same(a.permute(0, 2, 3, 1), packing.F1_DOWN, (33, 33), 2)
    us = E.space_to_depth(torch.nn.functional.pad(ub, [0, 0, 0, 1, 0, 1]))
    gotd = E.conv_taps_table(us, packing.exact_downconv(wd), packing.DOWN_EXACT_TAPS, (16, 16)).permute(0, 3, 1, 2)
    np.testing.assert_allclose(gotd.numpy(), refd.numpy(), atol=1e-4)


def test_skip_upsample_polyphase_matches_reference_upsample():
    from oracle import stylegan2_oracle as so
    y = torch.randn(2, 3, 6, 6)
    ref = so.upsample_skip(y)
    got = E.skip_upsample(y.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
    np.testing.assert_allclose(got.numpy(), ref.numpy(), atol=1e-6)


def test_style_offsets_and_layer_table():
    layers = packing.g_layers(W.FFHQ)
    assert len(layers) == 17 and W.FFHQ.num_noise_layers == 17 and W.FFHQ.num_style_layers == 18
    assert [l["res"] for l in layers][-2:] == [1024, 1024]
    conv_off, rgb_off, total = packing.style_offsets(W.FFHQ)
    assert total == sum(l["cin"] for l in layers) + sum(W.FFHQ.channels)
    assert sum(s * s for s in W.FFHQ.noise_shapes()) == 16 + 2 * sum((4 * 2 ** i) ** 2 for i in range(1, 9))


# ---------------------------------------------------------------------------
# C ABI surface
# ---------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(REPO, "include", "clipglass_b200.h")).read()
    declared = set(re.findall(r"\b(glass_[a-z0-9_]+)\s*\(", header))
    declared -= {"glass_engine", "glass_status", "glass_config", "glass_noise"}
    lib = _lib.load_library()
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert ctypes.sizeof(_lib.GlassConfig) == 4 * (1 + 12 + 14)
    assert ctypes.sizeof(_lib.GlassNoise) == 32


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_engine_creation_fails_loudly():
    from clip_glass_b200.engine import GlassEngine
    gan, clip = W.TINY_GAN, W.TINY_CLIP
    with pytest.raises(_lib.GlassError, match="no usable CUDA device|no CPU fallback"):
        GlassEngine(gan, clip, W.make_generator_weights(gan, 0), None, W.make_clip_visual_weights(clip, 2),
                    batch_size=4, max_population=8)


def test_bad_config_is_rejected_before_touching_the_gpu():
    lib = _lib.load_library()
    cfg = _lib.GlassConfig()
    cfg.num_blocks = 1
    h = ctypes.c_void_p()
    assert lib.glass_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"num_blocks" in lib.glass_last_error()


# ---------------------------------------------------------------------------
# host mirror of config.py / operators.py / latent.py
# ---------------------------------------------------------------------------
def test_config_values_match_reference_config_py():
    c = cfgmod.get_config("StyleGAN2_ffhq_d")       # config.py:74-94
    assert (c["pop_size"], c["batch_size"], c["algorithm"], c["dim_z"], c["task"]) == (16, 4, "nsga2", 512, "txt2img")
    assert c["problem_args"] == dict(n_var=512, n_obj=2, n_constr=512, xl=-10, xu=10)
    assert c["use_discriminator"] is True and c["weights"] == "./stylegan2/weights/ffhq-config-f"
    n = cfgmod.get_config("StyleGAN2_ffhq_nod")     # config.py:136-155
    assert (n["algorithm"], n["problem_args"]["n_obj"], n["use_discriminator"]) == ("ga", 1, False)
    c["pop_size"] = 64
    assert cfgmod.get_config("StyleGAN2_ffhq_d")["pop_size"] == 16     # copies, not shared state
    g = cfgmod.get_config("GPT2")                   # config.py:5-25
    assert (g["task"], g["dim_z"], g["max_tokens_len"], g["max_text_len"], g["encoder_size"], g["init_text"],
            g["stochastic"], g["pop_size"], g["batch_size"]) == ("img2txt", 20, 30, 50, 50257, "the picture of", False, 100, 25)
    assert g["problem_args"] == dict(n_var=20, n_obj=1, n_constr=20, xl=0, xu=50256)
    b = cfgmod.get_config("DeepMindBigGAN512")      # config.py:52-72
    assert (b["dim_z"], b["num_classes"], b["pop_size"], b["batch_size"], b["truncation"]) == (128, 1000, 32, 8, 1.0)
    assert b["problem_args"] == dict(n_var=1128, n_obj=1, n_constr=128, xl=-2, xu=2)
    x = torch.rand(2, 3, 4, 4) * 4 - 2
    np.testing.assert_allclose(c["norm"](x).numpy(), ((x + 1) / 2).clamp(0, 1).numpy())
    np.testing.assert_allclose(c["denorm"](x).numpy(), (x * 2 - 1).numpy())


def test_sampling_operators():
    class P:
        n_var = 512
    np.random.seed(0)
    x = operators.NormalRandomSampling()._do(P, 16)                 # operators.py:24-25
    assert x.shape == (16, 512) and x.dtype == np.float64 and abs(x.std() - 1) < 0.05
    np.random.seed(0)
    assert np.array_equal(x, np.random.normal(0, 1, size=(16, 512)))
    t = operators.TruncatedNormalRandomSampling()._do(P, 8)         # operators.py:14-15
    assert t.dtype == np.float32 and np.abs(t).max() <= 2.0
    b = operators.BinaryRandomSampling(prob=5 / 1000)._do(P, 64)    # operators.py:32-34
    assert b.dtype == bool and b.mean() < 0.02


def test_latent_space_surface():
    from clip_glass_b200.latent import StyleGAN2LatentSpace
    ns = cfgmod.make_namespace("StyleGAN2_ffhq_d", device="cpu")
    ls = StyleGAN2LatentSpace(ns)
    x = np.random.default_rng(0).normal(size=(8, 512))
    ls.set_from_population(x)
    (z,) = ls()
    assert z.dtype == torch.float32 and z.shape == (8, 512)
    np.testing.assert_array_equal(z.numpy(), x.astype(np.float32))
    assert "z" in ls.state_dict()                                   # run.py:101


# ---------------------------------------------------------------------------
# population sharding, world_size 2 over gloo
# ---------------------------------------------------------------------------
def test_shard_bounds_keep_minibatches_whole():
    for pop, b, w in [(512, 4, 8), (64, 4, 8), (12, 4, 2), (8, 4, 4), (100, 25, 3)]:
        bounds = dist.shard_bounds(pop, b, w)
        assert bounds[0][0] == 0 and bounds[-1][1] == pop
        for (s, e), (s2, _) in zip(bounds, bounds[1:] + [(pop, pop)]):
            assert e == s2 and s % b == 0 and e % b == 0
    with pytest.raises(AssertionError):
        dist.shard_bounds(10, 4, 2)


def _fake_eval(xs, first_group):
    """deterministic per-candidate stand-in for the GPU evaluation"""
    return -np.tanh(xs[:, :8].sum(1)).astype(np.float32), np.abs(xs[:, 8:16]).mean(1).astype(np.float32)


def _gloo_worker(rank, world, port, pop, n_obj, as_tensor, q):
    import torch.distributed as tdist
    tdist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    x = np.random.default_rng(3).normal(size=(pop, 32))
    groups = []

    def ev(xs, first_group):
        groups.append(first_group)
        a, b = _fake_eval(xs, first_group)
        if as_tensor:                      # the NCCL route hands device tensors to the gather; same code path
            a, b = torch.from_numpy(a), torch.from_numpy(b)
        return a, (b if n_obj == 2 else None)
    neg_sim, hinge = dist.sharded_evaluate(x, 4, n_obj, ev)
    q.put((rank, neg_sim, hinge, groups))
    tdist.destroy_process_group()


@pytest.mark.parametrize("pop,n_obj,as_tensor", [(16, 2, False), (12, 2, True), (4, 2, False), (12, 1, True), (4, 1, False)])
def test_sharded_evaluate_equals_single_rank_gloo(pop, n_obj, as_tensor):
    """World size 2 over gloo: ONE all-gather, no other collective (the column count comes from n_obj); uneven and
    empty shards; host-array and tensor outputs of the local evaluation."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + pop + 37 * n_obj + int(as_tensor)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, pop, n_obj, as_tensor, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    x = np.random.default_rng(3).normal(size=(pop, 32))
    exp_sim, exp_hinge = _fake_eval(x, 0)
    for rank, neg_sim, hinge, groups in res:
        np.testing.assert_array_equal(neg_sim, exp_sim)
        if n_obj == 2:
            np.testing.assert_array_equal(hinge, exp_hinge)
        else:
            assert hinge is None
        bounds = dist.shard_bounds(pop, 4, 2)
        if bounds[rank][1] > bounds[rank][0]:
            assert groups == [bounds[rank][0] // 4]
        else:
            assert groups == []


# ---------------------------------------------------------------------------
# evidence consistency: the summaries DESIGN.md / bench.py cite are what profiles/summarize.py derives from the
# committed ncu logs
def test_profile_summaries_match_the_committed_ncu_logs(tmp_path):
    import importlib.util
    import json
    prof = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
    spec = importlib.util.spec_from_file_location("summarize", os.path.join(prof, "summarize.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = tmp_path / "traffic.json"
    mod.traffic(os.path.join(prof, "r02_conv_dram_p64.csv"), str(out))
    got = json.load(open(out))
    ref = json.load(open(os.path.join(prof, "conv_tc_traffic.json")))
    assert got["launches"] == ref["launches"] == 92                    # one step = 92 tcgen05 conv/GEMM launches
    assert abs(got["bytes_per_launch"] - ref["bytes_per_launch"]) < 1.0
    shares = tmp_path / "shares.txt"
    mod.launches(os.path.join(prof, "r02_launches_p64.csv"), str(shares))
    assert open(shares).read() == open(os.path.join(prof, "r02_launch_shares_p64.txt")).read()
    met = tmp_path / "metrics.txt"
    mod.metrics(os.path.join(prof, "r02_metrics_p64.csv"), str(met))
    assert open(met).read() == open(os.path.join(prof, "r02_metrics_p64.txt")).read()


def test_python_flag_constants_match_the_header():
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include",
                            "clipglass_b200.h")).read()
    defs = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define GLASS_FLAG_(\w+)\s+(\d+)", hdr)}
    assert defs, "no GLASS_FLAG_* in the header"
    for name, value in defs.items():
        assert getattr(_lib, "FLAG_" + name) == value, name
    assert len(set(defs.values())) == len(defs) and all(v & (v - 1) == 0 for v in defs.values())   # distinct bits


# ---------------------------------------------------------------------------
# real-checkpoint loading: the reference's G.pth / D.pth pickle layout
# ---------------------------------------------------------------------------
def _reference_layout_blob(spec, seed):
    """What stylegan2.models.Generator._serialize() writes (stylegan2/models.py:111-132, 249-262): weights nested under
    'G_mapping' / 'G_synthesis', the top-level state_dict holding only the dlatent_avg buffer."""
    g_sd = W.make_generator_weights(spec, seed)
    sub = lambda pre: {k[len(pre):]: v for k, v in g_sd.items() if k.startswith(pre)}
    blob = dict(name="Generator", kwargs={}, state_dict={"dlatent_avg": torch.zeros(spec.latent_size)},
                G_mapping=dict(name="GeneratorMapping", state_dict=sub("G_mapping."),
                               kwargs=dict(latent_size=spec.latent_size, num_layers=spec.mapping_layers, lr_mul=0.01)),
                G_synthesis=dict(name="GeneratorSynthesis", state_dict=sub("G_synthesis."),
                                 kwargs=dict(channels=list(spec.channels), latent_size=spec.latent_size)))
    d_blob = dict(name="Discriminator", kwargs=dict(channels=list(spec.channels), mbstd_group_size=spec.mbstd_group_size),
                  state_dict=W.make_discriminator_weights(spec, seed + 1))
    return g_sd, blob, d_blob


@pytest.mark.parametrize("spec", [W.TINY_GAN, W.GanSpec(channels=(64, 128, 256, 512, 512, 512, 512), mbstd_group_size=4)])
def test_reference_checkpoint_layout_loads_and_packs(tmp_path, spec):
    """G.pth in the reference's nested layout -> flat state dict -> GanSpec from the checkpoint (not a hand-written
    default: the second case is the 256x256 church-config-f shape) -> packing, identical to packing the flat dict."""
    g_sd, blob, d_blob = _reference_layout_blob(spec, 11)
    torch.save(blob, tmp_path / "G.pth")
    torch.save(d_blob, tmp_path / "D.pth")
    flat, kw = W.load_reference_checkpoint(str(tmp_path / "G.pth"))
    assert "dlatent_avg" in flat and set(g_sd) <= set(flat)
    d_flat, d_kw = W.load_reference_checkpoint(str(tmp_path / "D.pth"))
    got = W.gan_spec_from_checkpoint(flat, kw, d_kw)
    assert got == spec
    a, b = packing.pack_generator(flat, got), packing.pack_generator(g_sd, spec)
    assert set(a) == set(b)
    for k in a:
        np.testing.assert_array_equal(np.asarray(a[k]), np.asarray(b[k]))
    packing.pack_discriminator(d_flat, got)
    # the generator.py entry point takes the same route
    from clip_glass_b200 import generator as gen
    ns = cfgmod.make_namespace("StyleGAN2_church_d", weights=str(tmp_path), clip_weights=str(tmp_path / "missing.pt"))
    with pytest.raises(Exception) as ei:       # gets past G/D loading + spec inference, stops at the absent CLIP archive
        gen._load_state_dicts(ns)
    assert "missing.pt" in str(ei.value) or "No such file" in str(ei.value) or "open file" in str(ei.value)


def test_unsupported_checkpoint_architecture_is_reported():
    g_sd, blob, d_blob = _reference_layout_blob(W.TINY_GAN, 12)
    blob["G_synthesis"]["kwargs"]["resnet"] = True
    flat, kw = W.flatten_reference_blob(blob)
    with pytest.raises(W.UnsupportedArchitecture, match="skip architecture"):
        W.gan_spec_from_checkpoint(flat, kw)


@pytest.mark.skipif(not os.path.isdir("/root/reference/stylegan2"), reason="needs the reference tree (build container)")
def test_reference_serialize_round_trip(tmp_path):
    """The real thing: the reference's own Generator / Discriminator ._serialize() -> torch.save -> our loader."""
    import sys
    sys.path.insert(0, "/root/reference")
    try:
        from stylegan2 import models
    finally:
        sys.path.remove("/root/reference")
    spec = W.TINY_GAN
    G = models.Generator(G_mapping=models.GeneratorMapping(latent_size=512, num_layers=8, lr_mul=0.01),
                         G_synthesis=models.GeneratorSynthesis(channels=list(spec.channels), latent_size=512))
    D = models.Discriminator(channels=list(spec.channels), mbstd_group_size=4)
    G.save(str(tmp_path / "G.pth"))
    D.save(str(tmp_path / "D.pth"))
    flat, kw = W.load_reference_checkpoint(str(tmp_path / "G.pth"))
    d_flat, d_kw = W.load_reference_checkpoint(str(tmp_path / "D.pth"))
    assert W.gan_spec_from_checkpoint(flat, kw, d_kw) == spec
    ref_sd = G.state_dict()
    for k, v in ref_sd.items():
        if k in flat:
            assert torch.equal(flat[k], v.float()), k
    assert all(k in flat for k, _ in G.named_parameters())
    packing.pack_generator(flat, spec)
    packing.pack_discriminator(d_flat, spec)


# ---------------------------------------------------------------------------
# GA operators (clip_glass_b200/ga.py; pymoo absent => parity unpinned, properties only) and operator wiring
# ---------------------------------------------------------------------------
class _Prob:
    def __init__(self, n_var, xl, xu):
        self.n_var, self.xl, self.xu = n_var, np.full(n_var, float(xl)), np.full(n_var, float(xu))


def test_sbx_and_pm_properties():
    from clip_glass_b200 import ga
    rng = np.random.RandomState(0)
    prob = _Prob(512, -10, 10)
    X = rng.normal(size=(2, 400, 512))
    sbx = ga.SimulatedBinaryCrossover(eta=3.0, prob=1.0, rng=rng)
    C = sbx.do(prob, X)
    assert C.shape == X.shape and (C >= -10).all() and (C <= 10).all()
    # children are symmetric around the parents' mean up to the bound correction of the spread factor (the two
    # children use the distance to their own bound), which only matters in the far tail of the spread distribution
    asym = np.abs(C.sum(0) - X.sum(0))
    assert np.median(asym) < 1e-4 and np.quantile(asym, 0.95) < 5e-2
    changed = (C[0] != X[0]) & (C[0] != X[1])
    assert 0.4 < changed.mean() < 0.6                      # prob_per_variable = 0.5
    same = np.stack([X[0], X[0]])
    np.testing.assert_array_equal(sbx.do(prob, same), same)      # identical parents are returned unchanged
    # larger distribution index => children closer to the parents
    spread = lambda eta: np.abs(ga.SimulatedBinaryCrossover(eta, 1.0, rng=np.random.RandomState(1)).do(prob, X)[0]
                                - X.mean(0)).mean()
    assert spread(30.0) < spread(3.0) * 1.05 and spread(3.0) > 0
    pm = ga.PolynomialMutation(eta=3.0, prob=0.5, rng=rng)
    Y = pm.do(prob, X[0])
    assert (Y >= -10).all() and (Y <= 10).all()
    assert 0.45 < (Y != X[0]).mean() < 0.55
    edge = np.full((50, 512), 10.0)
    assert (pm.do(prob, edge) <= 10).all()


def test_integer_and_binary_operators():
    from clip_glass_b200 import ga
    rng = np.random.RandomState(2)
    prob = _Prob(20, 0, 50256)
    S = ga.IntegerRandomSampling(rng)._do(prob, 64)
    assert S.shape == (64, 20) and S.min() >= 0 and S.max() <= 50256 and S.dtype.kind == "i"
    X = np.stack([S[:32], S[32:]])
    C = ga.get_crossover("int_sbx", prob=1.0, eta=3.0).do(prob, X)
    M = ga.get_mutation("int_pm", prob=0.5, eta=3.0).do(prob, C.reshape(-1, 20))
    for arr in (C, M):
        assert arr.dtype.kind == "i" and arr.min() >= 0 and arr.max() <= 50256
    B = rng.random((2, 30, 1000)) < 0.005
    H = ga.HalfUniformCrossover(prob=1.0, rng=rng).do(_Prob(1000, 0, 1), B)
    np.testing.assert_array_equal(H.sum(0), B.sum(0))                 # genes are exchanged, never created
    diff = (B[0] != B[1]).sum(1)
    np.testing.assert_array_equal((H[0] != B[0]).sum(1), np.ceil(diff / 2).astype(int))
    F = ga.BitflipMutation(prob=0.01, rng=rng).do(_Prob(1000, 0, 1), B[0])
    assert 0.005 < (F != B[0]).mean() < 0.015


def test_non_dominated_sort_and_crowding():
    from clip_glass_b200 import ga
    F = np.array([[0, 5], [1, 3], [2, 2], [4, 1], [3, 3], [5, 5], [1, 4], [6, 6]], dtype=float)
    fronts = ga.fast_non_dominated_sort(F)
    assert sorted(fronts[0]) == [0, 1, 2, 3] and sorted(fronts[1]) == [4, 6] and sorted(fronts[2]) == [5]
    cd = ga.crowding_distance(F[fronts[0]])
    assert np.isinf(cd[[0, 3]]).all() and np.isfinite(cd[[1, 2]]).all()
    idx, rank, _ = ga.rank_and_crowding_survival(F, 5)
    assert set(idx[:4]) == {0, 1, 2, 3} and rank[4] == 1


def test_get_operators_wires_every_config():
    """operators.py:37-82: all three branches (StyleGAN2 real, BigGAN mixed real/bool, GPT-2 integer)."""
    for name in ("StyleGAN2_ffhq_d", "DeepMindBigGAN512", "GPT2"):
        ns = cfgmod.make_namespace(name, device="cpu")
        ops = operators.get_operators(ns)
        prob = _Prob(ns.problem_args["n_var"], ns.problem_args["xl"], ns.problem_args["xu"])
        X = ops["sampling"]._do(prob, 8)
        assert X.shape == (8, ns.problem_args["n_var"])
        C = ops["crossover"].do(prob, np.stack([X[:4], X[4:]]))
        Y = ops["mutation"].do(prob, C.reshape(8, -1))
        assert Y.shape == X.shape
        if name == "DeepMindBigGAN512":
            zf = Y[:, :128].astype(float)
            assert np.abs(zf).max() <= 2.0 and set(np.unique(Y[:, 128:].astype(int))) <= {0, 1}
            ls = ns.latent(ns)
            ls.set_from_population(Y)
            z, cl = ls()                                  # latent.py:20-24 on the CPU route (torch ops)
            assert z.shape == (8, 128) and cl.shape == (8, 1000) and float(z.abs().max()) <= 2.0
            np.testing.assert_allclose(cl.sum(1).numpy(), 1.0, rtol=1e-5)
        if name == "GPT2":
            assert Y.dtype.kind == "i" and Y.min() >= 0 and Y.max() <= 50256
    with pytest.raises(Exception, match="Unknown config"):
        operators.get_operators(cfgmod.Namespace(config="nope"))


def test_ga_loop_minimises_a_toy_problem():
    """The stand-in for pymoo's minimize (run.py:59-76): GA on a sphere, NSGA-II on a two-objective toy."""
    from clip_glass_b200 import ga

    class Toy(_Prob):
        def __init__(self, n_obj):
            super().__init__(6, -10, 10)
            self.n_obj, self.calls = n_obj, []

        def _evaluate(self, x, out):
            self.calls.append(x.shape[0])
            f1 = (x ** 2).sum(1)
            out["F"] = f1 if self.n_obj == 1 else np.column_stack((f1, ((x - 2) ** 2).sum(1)))

    for n_obj, name in ((1, "ga"), (2, "nsga2")):
        toy = Toy(n_obj)
        alg = ga.get_algorithm(name, pop_size=16, sampling=operators.NormalRandomSampling(),
                               crossover=ga.get_crossover("real_sbx", prob=1.0, eta=3.0),
                               mutation=ga.get_mutation("real_pm", prob=0.5, eta=3.0), seed=0)
        first = []
        alg.callback = lambda a: first.append(np.min(a.pop.get("F").reshape(len(a.pop), -1)[:, 0]))
        res = ga.minimize(toy, alg, ("n_gen", 30), seed=0)
        assert first[-1] < first[0] and all(c == 16 for c in toy.calls)
        assert len(res.pop) == 16
        if n_obj == 2:
            assert len(ga.fast_non_dominated_sort(np.atleast_2d(res.F))[0]) == len(np.atleast_2d(res.F))


# ---------------------------------------------------------------------------
# img2txt host side: tokenizers, parse_out, oracle vs fixtures
# ---------------------------------------------------------------------------
_GPT2_VOCAB = ("/root/reference/gpt2/weights/encoder.json", "/root/reference/gpt2/weights/vocab.bpe")


@pytest.mark.skipif(not all(os.path.exists(p) for p in _GPT2_VOCAB), reason="needs the GPT-2 vocabulary of the reference tree")
def test_gpt2_tokenizer_matches_reference_encoder():
    """clip_glass_b200.tokenizers.GPT2Tokenizer against the reference's gpt2/encoder.py on its own vocabulary:
    encode and decode of text, and decode of arbitrary token lists (what parse_out sees)."""
    import sys
    from clip_glass_b200.tokenizers import GPT2Tokenizer
    sys.path.insert(0, "/root/reference")
    try:
        from gpt2.encoder import get_encoder
    finally:
        sys.path.remove("/root/reference")
    ref = get_encoder(cfgmod.Namespace(encoder=_GPT2_VOCAB[0], vocab=_GPT2_VOCAB[1]))
    mine = GPT2Tokenizer(*_GPT2_VOCAB)
    assert mine.encode("the picture of") == ref.encode("the picture of") == [1169, 4286, 286]
    texts = ["the picture of a wolf at night with the moon in the background", "Hello, world!  It's 2021... isn't it?",
             "naïve café — “quotes” and emoji 🙂", "   leading and trailing   ", "tabs\tand\nnewlines\n\n", "a", ""]
    for t in texts:
        assert mine.encode(t) == ref.encode(t), t
        assert mine.decode(mine.encode(t)) == t
    rng = np.random.default_rng(0)
    for _ in range(50):
        ids = rng.integers(0, 50257, size=int(rng.integers(1, 40))).tolist()
        assert mine.decode(ids) == ref.decode(ids)
    # parse_out (models.py:32-42): cut at the first EOT anywhere in the sequence, decode, truncate to 50 characters
    seqs = rng.integers(0, 50256, size=(4, 53))
    seqs[1, 30] = 50256          # generated EOT
    seqs[2, 5] = 50256           # EOT gene inside the latent part -> empty text
    got = mine.parse_out(seqs, 20, 50)
    assert got[0] == ref.decode(seqs[0, 20:].tolist())[:50] and got[1] == ref.decode(seqs[1, 20:30].tolist())[:50]
    assert got[2] == "" and all(len(t) <= 50 for t in got)


def test_clip_tokenizer_shape_and_errors():
    """CLIP's BPE vocabulary is not in the reference tree: the tokenizer is exercised on a synthetic merge table
    (parity unpinned); tokenize() keeps the reference's contract (SOT/EOT/zero padding, RuntimeError when too long)."""
    from clip_glass_b200.tokenizers import ClipTokenizer, byte_alphabet
    b = byte_alphabet()
    assert len(set(b.values())) == 256 and b[ord("a")] == "a" and b[32] == chr(256 + 32)
    tok = ClipTokenizer(merges=[("t", "h"), ("th", "e</w>"), ("c", "a"), ("ca", "t</w>")])
    ids = tok.encode("The  CAT the")
    assert ids == [tok.encoder["the</w>"], tok.encoder["cat</w>"], tok.encoder["the</w>"]]
    assert tok.decode(ids) == "the cat the "
    t = tok.tokenize(["the cat", "cat"], context_length=8)
    assert t.shape == (2, 8) and t[0, 0] == tok.sot and t[0, 3] == tok.eot and t[0, 4:].sum() == 0
    assert (t.argmax(1) == [3, 2]).all()                      # the EOT id is the largest: clip/model.py:318 gathers it
    with pytest.raises(RuntimeError):
        tok.tokenize(["the " * 10], context_length=8)


@pytest.mark.parametrize("name", ["gpt2_tiny", "gpt2_full"])
def test_gpt2_oracle_reproduces_reference_fixture(name):
    """oracle/gpt2_oracle.py against tests/golden/gpt2_*.npz (reference gpt2.model + sample_sequence, clip.model):
    tokens bit-exact, text cosine in fp16-as-built mode within fp16 rounding of the reference's."""
    from clip_glass_b200 import text_weights as TW
    from oracle import gpt2_oracle
    from tests.test_gpu_text import CONFIGS
    cfg = CONFIGS[name]
    gold = dict(np.load(os.path.join(REPO, "tests", "golden", f"{name}.npz")))
    if name == "gpt2_full":
        z, want = gold["z"][:2], gold["tokens"][:2]           # keep the CPU suite short
    else:
        z, want = gold["z"], gold["tokens"]
    got = gpt2_oracle.gpt2_generate_tokens(TW.make_gpt2_weights(cfg["gpt2"], cfg["seed"]), cfg["gpt2"], z,
                                           gold["init_tokens"].tolist(), 30)
    np.testing.assert_array_equal(got, want)
    assert gpt2_oracle.parse_out_tokens(gold["tokens"], 20, cfg["gpt2"].vocab - 1)[1] == []
    built = TW.text_as_built(TW.make_clip_text_weights(cfg["text"], cfg["seed"] + 1))
    n = 8 if name == "gpt2_tiny" else 2
    feats = gpt2_oracle.clip_encode_text(built, cfg["text"], torch.tensor(gold["clip_tokens"][:n]), mode="fp32")
    sim = gpt2_oracle.text_similarity(feats, torch.from_numpy(gold["image_features"])).numpy()
    np.testing.assert_allclose(sim, gold["sim_oracle_fp32"][:n], rtol=1e-5)
    np.testing.assert_allclose(sim, gold["sim_fp16"][:n].astype(np.float32), rtol=3e-3)


# ---------------------------------------------------------------------------
# driver plumbing for population-sharded runs
# ---------------------------------------------------------------------------
def test_generator_renders_more_candidates_than_the_workspace_in_chunks():
    """A population-sharded run sizes each engine for its shard; the saving rank still renders the whole population
    (run.py:45): whole minibatches per engine call, a fresh noise seed per call, rows in order."""
    from types import SimpleNamespace
    from clip_glass_b200.generator import Generator

    calls = []

    class Eng:
        max_population = 8

        def set_batch_size(self, b):
            self.b = b

        def generate(self, z, noise=None, seed=0):
            assert z.shape[0] <= self.max_population and z.shape[0] % self.b == 0 and z.is_contiguous()
            calls.append((z.shape[0], seed, self.b))
            return z[:, :3, None, None].expand(-1, 3, 2, 2).clone()

    g = Generator.__new__(Generator)
    g.engine, g.config, g._calls = Eng(), SimpleNamespace(noise_seed=10), 0
    z = torch.arange(20 * 4, dtype=torch.float32).reshape(20, 4)
    out = g._render(z, 4)
    assert calls == [(8, 11, 4), (8, 12, 4), (4, 13, 4)] and out.shape == (20, 3, 2, 2)
    assert torch.equal(out[:, :, 0, 0], z[:, :3])
    calls.clear()
    out = g._render(z[:8], 8)                              # fits: one call, exactly as before
    assert calls == [(8, 14, 8)] and torch.equal(out[:, :, 0, 0], z[:8, :3])
    with pytest.raises(_lib.GlassArgError):
        g._render(z, 16)                                   # a minibatch larger than the workspace
    with pytest.raises(_lib.GlassArgError):
        g._render(z, 4, noise=[object()])


def _init_env_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as tdist
    got = dist.init_from_env("cpu", timeout_s=45.0)
    again = dist.init_from_env("cpu")                       # idempotent
    x = np.random.default_rng(3).normal(size=(12, 32))
    neg_sim, hinge = dist.sharded_evaluate(x, 4, 2, _fake_eval)
    q.put((rank, got, again, tdist.get_backend(), neg_sim, hinge))
    tdist.destroy_process_group()


def test_driver_creates_the_process_group_from_the_torchrun_environment():
    import torch.multiprocessing as mp
    assert dist.init_from_env("cpu") == (0, 1, 0) and not dist.tdist.is_initialized()      # plain launch: nothing
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_init_env_worker, args=(r, 2, 29871, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    x = np.random.default_rng(3).normal(size=(12, 32))
    exp_sim, exp_hinge = _fake_eval(x, 0)
    for r, (rank, got, again, backend, neg_sim, hinge) in enumerate(res):
        assert got == (r, 2, r) and again == got and backend == "gloo"
        np.testing.assert_array_equal(neg_sim, exp_sim)
        np.testing.assert_array_equal(hinge, exp_hinge)


def test_ga_entry_points_validate_arguments_without_a_gpu():
    """glass_ga_* check their arguments before touching CUDA (GLASS_ERR_ARG = -1 with a message); the size helpers
    are pure host arithmetic."""
    lib = _lib.load_library()
    assert lib.glass_ga_rand_count(32, 512) == 7 * 32 * 512 + 32
    assert lib.glass_ga_dedup_workspace(64) == 64 * 8
    n = 128
    assert lib.glass_ga_survive_workspace(n) >= 4 * n * 4 + 2 * n * 8
    one = ctypes.c_void_p(16)                 # never dereferenced: every call below is refused first
    assert lib.glass_ga_permutations(one, 5000, 1, one, None) == -1
    assert b"4096" in lib.glass_ga_last_error()
    assert lib.glass_ga_survive(one, 4, 8, 2, 4, 1, one, one, one, one, None) == -1       # ld < n
    assert lib.glass_ga_survive(one, 8, 8, 2, 9, 1, one, one, one, one, None) == -1       # n_survive > n
    assert lib.glass_ga_survive(one, 8, 8, 9, 4, 1, one, one, one, one, None) == -1       # n_obj > 8
    assert lib.glass_ga_uniform(0, 0, None, 4, None) == -1
    assert lib.glass_ga_tournament(one, one, one, 0, one, None) == -1
    p = _lib.GlassGaParams(3.0, 1.0, 0.5, 3.0, 0.5, 0, 0)
    assert lib.glass_ga_offspring(ctypes.byref(p), one, one, one, one, 4, one, None) == -1  # n_var = 0
    assert lib.glass_ga_gather(one, one, 8, one, 4, 2, 3, one, one, 8, None) == -1          # n_obj > n_var
    with pytest.raises(_lib.GlassArgError):
        _lib.check_ga(lib.glass_ga_pad(None, 4, one, 4, None, None))
    assert ctypes.sizeof(_lib.GlassGaParams) == 5 * 8 + 2 * 4


def test_generator_generate_branches_with_a_stub_engine():
    """Generator.generate (generator.py:29-34 + the image reuse of SURVEY §8(f)-4) over a stub engine: nothing cached
    -> one engine.generate call; everything cached -> images gathered from the last evaluation, no rendering; partly
    cached -> the missing rows rendered as whole minibatches (padded by repeating the last missing row)."""
    from types import SimpleNamespace
    from clip_glass_b200.generator import Generator
    log = []

    class Eng:
        max_population = 8

        def set_batch_size(self, b):
            log.append(("batch", b))

        def generate(self, z, noise=None, seed=0):
            log.append(("generate", z.shape[0], seed))
            return z[:, :3, None, None].expand(-1, 3, 2, 2).clone()

        def last_images(self, rows):
            log.append(("last_images", list(rows)))
            return torch.full((len(rows), 3, 2, 2), -1.0)

    g = Generator.__new__(Generator)
    g.engine, g._calls = Eng(), 0
    g.config = SimpleNamespace(task="txt2img", device="cpu", noise_seed=100)
    g.gan = SimpleNamespace(resolution=2)
    x = np.random.default_rng(0).normal(size=(8, 4))
    ls = lambda rows: (lambda: (torch.from_numpy(x[rows]).float(),))
    out = g.generate(ls(slice(0, 8)), minibatch=4)                         # nothing remembered yet
    assert log == [("batch", 4), ("generate", 8, 101)] and out.shape == (8, 3, 2, 2)
    log.clear()
    g.remember_population(x)
    out = g.generate(ls([5, 2]), minibatch=4)                              # both rows were scored by the last evaluation
    assert log == [("last_images", [5, 2])] and float(out.max()) == -1.0 and g.reuse_stats == dict(reused=2, rendered=0)
    log.clear()
    g.remember_population(x[:4])
    out = g.generate(ls([1, 6, 3, 7, 5]), minibatch=2)                     # rows 6, 7, 5 are new: 3 -> padded to 4
    assert log == [("last_images", [1, 3]), ("batch", 2), ("generate", 4, 102)]
    assert torch.equal(out[[0, 2]], torch.full((2, 3, 2, 2), -1.0))
    assert torch.equal(out[[1, 3, 4], :, 0, 0], torch.from_numpy(x[[6, 7, 5], :3]).float())
    assert g.reuse_stats == dict(reused=2, rendered=3)
    log.clear()
    g.forget_population()
    g.generate(ls([0]))                                                     # run.py:117-118: one candidate, minibatch=None
    assert log == [("batch", 1), ("generate", 1, 103)]
    out = g.generate(ls(slice(0, 8)), minibatch=4, noise=[["explicit"]])   # explicit noise never reuses cached images
    assert log[-1] == ("generate", 8, 104)


def test_generation_problem_evaluate_contract_with_a_stub_engine():
    """problem.py:14-29 through the host mirror with the GPU engine stubbed out: F = (-sim, hinge) columns for the
    NSGA-II config, F = -sim for the GA config, G zeros, the noise seed advancing per generation, the population
    remembered for the image output path; pop % batch_size is the engine's assertion."""
    from types import SimpleNamespace
    from clip_glass_b200.problem import GenerationProblem
    seen = []

    class Eng:
        def set_batch_size(self, b):
            self.b = b

        def evaluate(self, xs, noise=None, seed=0, first_group=0):
            seen.append((xs.shape, seed, first_group, self.b))
            return (-xs[:, 0]).astype(np.float32), np.abs(xs[:, 1]).astype(np.float32)

    for n_obj, use_d in ((2, True), (1, False)):
        remembered = []
        gen = SimpleNamespace(engine=Eng(), remember_population=lambda xs: remembered.append(xs.shape))
        cfg = SimpleNamespace(task="txt2img", batch_size=4, noise_seed=7, use_discriminator=use_d, device="cuda:0",
                              problem_args=dict(n_var=6, n_obj=n_obj, n_constr=0, xl=-10.0, xu=10.0))
        p = GenerationProblem(cfg, generator=gen)
        x = np.random.default_rng(n_obj).normal(size=(8, 6))
        seen.clear()
        for g in (1, 2):
            out = {}
            p._evaluate(x, out)
            assert seen[-1] == ((8, 6), 7 + g, 0, 4) and remembered[-1] == (8, 6)
            if n_obj == 2:
                assert out["F"].shape == (8, 2)
                np.testing.assert_array_equal(out["F"][:, 0], (-x[:, 0]).astype(np.float32))
                np.testing.assert_array_equal(out["F"][:, 1], np.abs(x[:, 1]).astype(np.float32))
            else:
                np.testing.assert_array_equal(out["F"], (-x[:, 0]).astype(np.float32))
            assert out["G"].shape == (8,) and not out["G"].any()
        assert p.n_var == 6 and p.n_obj == n_obj and p.xl.shape == (6,)


class _StubGenerator:
    def __init__(self):
        self.saved, self.generated = [], []

    def generate(self, ls, minibatch=None):
        z = ls()[0]
        self.generated.append((tuple(z.shape), minibatch))
        return torch.zeros(z.shape[0], 3, 2, 2)

    def save(self, images, path):
        self.saved.append(os.path.basename(path))
        open(path, "wb").write(b"x")


class _StubProblem:
    """GenerationProblem's surface (problem.py:8-29) with an analytic fitness instead of the GPU engine."""

    def __init__(self, config):
        self.config, self.generator = config, _StubGenerator()
        a = config.problem_args
        self.n_var, self.n_obj = a["n_var"], a["n_obj"]
        self.xl, self.xu = np.full(self.n_var, float(a["xl"])), np.full(self.n_var, float(a["xu"]))

    def _evaluate(self, x, out, *args, **kwargs):
        f = (np.asarray(x, dtype=float) ** 2).sum(1)
        out["F"] = np.column_stack((f, np.abs(np.asarray(x, dtype=float)[:, 0] - 1))) if self.n_obj == 2 else f
        out["G"] = np.zeros(len(x))


_DRIVER_ARGS = ["--device", "cpu", "--generations", "5", "--save-each", "2", "--pop-size", "8", "--batch-size", "4",
                "--synthetic-seed", "1"]


@pytest.mark.parametrize("config_name", ["StyleGAN2_ffhq_nod", "StyleGAN2_ffhq_d"])
def test_run_driver_flow_with_a_stub_problem(tmp_path, monkeypatch, config_name):
    """The driver mirror (run.py:15-125) end to end on the CPU with the fitness engine stubbed out: sampling -> host
    GA / NSGA-II -> save_callback files -> result pickles -> final output; single process (no process group)."""
    import pickle
    from clip_glass_b200 import run as driver
    monkeypatch.setattr(driver, "GenerationProblem", _StubProblem)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    res = driver.main(_DRIVER_ARGS + ["--config", config_name, "--tmp-folder", str(tmp_path), "--seed", "3"])
    names = set(os.listdir(tmp_path))
    assert {"genetic-it-2.jpg", "genetic-it-4.jpg", "genetic-it-final.jpg", "output.jpg", "genetic_result",
            "ls_result"} <= names, names
    assert len(res.pop) == 8 and not dist.tdist.is_initialized()
    with open(tmp_path / "genetic_result", "rb") as f:
        saved = pickle.load(f)
    assert np.isfinite(np.asarray(saved["F"], dtype=float)).all()
    assert "z" in torch.load(tmp_path / "ls_result")


def _driver_worker(rank, world, port, folder, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from clip_glass_b200 import run as driver
    driver.GenerationProblem = _StubProblem
    res = driver.main(_DRIVER_ARGS + ["--config", "StyleGAN2_ffhq_d", "--tmp-folder", folder])      # no --seed given
    q.put((rank, np.stack([p.X for p in res.pop]), dist.tdist.is_initialized(), dist.tdist.get_backend()))
    dist.tdist.destroy_process_group()


def test_run_driver_under_a_two_rank_launch_writes_files_once(tmp_path):
    """torchrun-style launch (gloo, world size 2): the driver creates the process group, every rank runs the same
    seeded search (the missing --seed becomes 0 on all ranks), and only rank 0 writes the files."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    folders = [str(tmp_path / f"rank{r}") for r in range(2)]       # separate folders show who wrote what
    procs = [ctx.Process(target=_driver_worker, args=(r, 2, 29877, folders[r], q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    (_, X0, init0, backend0), (_, X1, init1, _) = res
    assert init0 and init1 and backend0 == "gloo"
    assert np.array_equal(X0, X1)                                  # same search on both ranks
    assert {"genetic-it-final.jpg", "output.jpg", "genetic_result", "ls_result"} <= set(os.listdir(folders[0]))
    assert os.listdir(folders[1]) == []                            # the folder is created, nothing is written
