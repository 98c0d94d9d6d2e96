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
    with pytest.raises(NotImplementedError):
        cfgmod.get_config("GPT2")
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


def _gloo_worker(rank, world, port, pop, q):
    import torch.distributed as tdist
    tdist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    x = np.random.default_rng(3).normal(size=(pop, 32))
    groups = []

    def ev(xs, first_group):
        groups.append(first_group)
        return _fake_eval(xs, first_group)
    neg_sim, hinge = dist.sharded_evaluate(x, 4, ev)
    q.put((rank, neg_sim, hinge, groups))
    tdist.destroy_process_group()


@pytest.mark.parametrize("pop", [16, 12, 4])
def test_sharded_evaluate_equals_single_rank_gloo(pop):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + pop
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, pop, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    x = np.random.default_rng(3).normal(size=(pop, 32))
    exp_sim, exp_hinge = _fake_eval(x, 0)
    for rank, neg_sim, hinge, groups in res:
        np.testing.assert_array_equal(neg_sim, exp_sim)
        np.testing.assert_array_equal(hinge, exp_hinge)
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
    mod.traffic(os.path.join(prof, "r01_conv_dram_p64_final.csv"), str(out))
    got = json.load(open(out))
    ref = json.load(open(os.path.join(prof, "conv_tc_traffic.json")))
    assert got["launches"] == ref["launches"] == 92                    # one step = 92 tcgen05 conv/GEMM launches
    assert abs(got["bytes_per_launch"] - ref["bytes_per_launch"]) < 1.0
    shares = tmp_path / "shares.txt"
    mod.launches(os.path.join(prof, "r01_launches_p64_final.csv"), str(shares))
    assert open(shares).read() == open(os.path.join(prof, "r01_launch_shares_p64_final.txt")).read()


def test_python_flag_constants_match_the_header():
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include",
                            "clipglass_b200.h")).read()
    defs = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define GLASS_FLAG_(\w+)\s+(\d+)", hdr)}
    assert defs, "no GLASS_FLAG_* in the header"
    for name, value in defs.items():
        assert getattr(_lib, "FLAG_" + name) == value, name
    assert len(set(defs.values())) == len(defs) and all(v & (v - 1) == 0 for v in defs.values())   # distinct bits
