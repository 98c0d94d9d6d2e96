"""-m gpu: the GPU-resident genetic operators (clip_glass_b200/csrc/ga.cu through the glass_ga_* C ABI) against the host
operators of clip_glass_b200/ga.py on the same uniform draws (cases: tests/ga_cases.py, also run on the CPU against the
host compilation of the same arithmetic in tests/test_ga_native.py), and the resident generation loop.

Tolerances: integer work (survivor indices, ranks, tournament winners, permutations, duplicate flags, the integer
operators) and the crowding distances are bit-exact; real-valued children are within 1e-13 relative of the host
operators (CUDA's pow and glibc's differ in the last bit; every other operation is the same IEEE operation in the same
order)."""
import ctypes
import os
import pickle

import numpy as np
import pytest
import torch

from clip_glass_b200 import weights as W
from clip_glass_b200._lib import GlassGaParams, check_ga, load_library
from tests import ga_cases as C

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


_KEEP = []


def dev(a):
    """Device copy that outlives the asynchronous launch it is passed to (only .data_ptr() goes through ctypes)."""
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    _KEEP.append(t)
    return t


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream(torch.device(DEV)).cuda_stream)


def test_uniform_stream_on_device():
    lib = load_library()
    n = 200001
    u = torch.empty(n, dtype=torch.float64, device=DEV)
    check_ga(lib.glass_ga_uniform(7, 0, u.data_ptr(), n, stream()))
    h = u.cpu().numpy()
    assert h.min() >= 0.0 and h.max() < 1.0 and len(np.unique(h)) == n
    assert abs(h.mean() - 0.5) < 4 * np.sqrt(1 / 12 / n) and abs(h.var() - 1 / 12) < 1e-3
    tail = torch.empty(n - 1000, dtype=torch.float64, device=DEV)
    check_ga(lib.glass_ga_uniform(7, 500, tail.data_ptr(), n - 1000, stream()))
    assert np.array_equal(tail.cpu().numpy(), h[1000:])
    # the host compilation of the same generator is pinned to the Random123 known-answer table in
    # tests/test_ga_native.py; the device build must produce the same stream bit for bit
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    so = os.path.join("/tmp", f"ga_host_{os.getpid()}.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                    os.path.join(here, "native", "ga_host.cpp"), "-o", so], check=True)
    host = ctypes.CDLL(so)
    ref = np.empty(n)
    host.ga_host_uniform(ctypes.c_uint64(7), ctypes.c_uint64(0), ref.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(n))
    assert np.array_equal(ref, h)
    os.remove(so)


def test_permutations_and_tournament_on_device():
    lib = load_library()
    rng = np.random.default_rng(0)
    keys = rng.random((5, 64))
    keys[2, 10] = keys[2, 3]
    out = torch.empty(5, 64, dtype=torch.int32, device=DEV)
    check_ga(lib.glass_ga_permutations(dev(keys).data_ptr(), 64, 5, out.data_ptr(), stream()))
    assert np.array_equal(out.cpu().numpy(), np.argsort(keys, axis=1, kind="stable"))
    big = rng.random((2, 4096))
    outb = torch.empty(2, 4096, dtype=torch.int32, device=DEV)
    check_ga(lib.glass_ga_permutations(dev(big).data_ptr(), 4096, 2, outb.data_ptr(), stream()))
    assert np.array_equal(outb.cpu().numpy(), np.argsort(big, axis=1, kind="stable"))
    for seed, n, n_select in ((0, 64, 64), (1, 10, 24), (2, 512, 512)):
        k = C.tournament_case(seed, n, n_select)
        sel = torch.empty(n_select, dtype=torch.int32, device=DEV)
        check_ga(lib.glass_ga_tournament(dev(k["pairs"]).data_ptr(), dev(k["rank"]).data_ptr(),
                                         dev(k["crowd"]).data_ptr(), n_select, sel.data_ptr(), stream()))
        assert np.array_equal(sel.cpu().numpy(), k["expect"])


@pytest.mark.parametrize("case", C.OFFSPRING_CASES, ids=lambda c: "int" if c["integer"] else f"real{c['V']}")
def test_offspring_kernel_against_host_operators(case):
    lib = load_library()
    # the uniforms come from the device generator, so the test covers the layout the resident loop uses
    n_rand = int(lib.glass_ga_rand_count(case["M"], case["V"]))
    assert n_rand == C.rand_count(case["M"], case["V"])
    rnd = torch.empty(n_rand, dtype=torch.float64, device=DEV)
    check_ga(lib.glass_ga_uniform(case["seed"], 0, rnd.data_ptr(), n_rand, stream()))
    k = C.offspring_case(**case, uniforms=rnd.cpu().numpy())
    p = GlassGaParams(**k["params"])
    out = torch.full((2 * k["M"], case["V"]), float("nan"), dtype=torch.float64, device=DEV)
    check_ga(lib.glass_ga_offspring(ctypes.byref(p), dev(k["X"]).data_ptr(), dev(k["parents"]).data_ptr(),
                                    dev(k["bounds"]).data_ptr(), rnd.data_ptr(), k["M"], out.data_ptr(), stream()))
    got = out.cpu().numpy()
    if case["integer"]:
        # rounding absorbs the last bit of pow unless a child sits within 1e-13 of a half-integer
        assert (got != k["expect"]).mean() < 1e-3 and np.abs(got - k["expect"]).max() <= 1.0
        assert np.array_equal(got, np.rint(got))
    else:
        np.testing.assert_allclose(got, k["expect"], rtol=1e-13, atol=1e-13)
    assert got.min() >= case["xl"] and got.max() <= case["xu"]


@pytest.mark.parametrize("have,n_off", [(0, 16), (3, 8), (2, 64)])
def test_duplicate_elimination_on_device(have, n_off):
    lib = load_library()
    k = C.dedup_case(seed=5, n_x=12, n_c=14, n_off=n_off, have=have, V=24)
    off, z32 = dev(k["off"]), torch.zeros(n_off, 24, dtype=torch.float32, device=DEV)
    n_have = torch.tensor([have], dtype=torch.int32, device=DEV)
    ws = torch.zeros(int(lib.glass_ga_dedup_workspace(14)), dtype=torch.uint8, device=DEV)
    check_ga(lib.glass_ga_dedup_append(dev(k["cand"]).data_ptr(), 14, dev(k["X"]).data_ptr(), 12, off.data_ptr(),
                                       n_off, n_have.data_ptr(), 24, 1e-16, 1, z32.data_ptr(), ws.data_ptr(),
                                       stream()))
    got = int(n_have.item())
    assert got == len(k["expect"])
    assert np.array_equal(off.cpu().numpy()[:got], k["expect"])
    assert np.array_equal(z32.cpu().numpy()[have:got], k["expect"][have:].astype(np.float32))
    # padding repeats the last accepted row
    check_ga(lib.glass_ga_pad(off.data_ptr(), n_off, n_have.data_ptr(), 24, z32.data_ptr(), stream()))
    full = off.cpu().numpy()
    assert np.array_equal(full[:got], k["expect"]) and all(np.array_equal(r, k["expect"][-1]) for r in full[got:])
    # eliminate = 0: every candidate is appended until the buffer is full
    off2, n2 = torch.zeros(n_off, 24, dtype=torch.float64, device=DEV), torch.zeros(1, dtype=torch.int32, device=DEV)
    check_ga(lib.glass_ga_dedup_append(dev(k["cand"]).data_ptr(), 14, dev(k["X"]).data_ptr(), 12, off2.data_ptr(),
                                       n_off, n2.data_ptr(), 24, 1e-16, 0, None, ws.data_ptr(), stream()))
    m = min(14, n_off)
    assert int(n2.item()) == m and np.array_equal(off2.cpu().numpy()[:m], k["cand"][:m])


@pytest.mark.parametrize("case", C.SURVIVE_CASES, ids=lambda c: f"n{c['n']}m{c['n_obj']}s{c['n_survive']}")
def test_survival_kernel_against_host(case):
    lib = load_library()
    k = C.survive_case(**case)
    S = k["n_survive"]
    idx = torch.full((S,), -1, dtype=torch.int32, device=DEV)
    rank = torch.full((S,), -1, dtype=torch.int32, device=DEV)
    crowd = torch.full((S,), float("nan"), dtype=torch.float64, device=DEV)
    ws = torch.zeros(int(lib.glass_ga_survive_workspace(k["n"])), dtype=torch.uint8, device=DEV)
    F = dev(k["Fcm"])
    check_ga(lib.glass_ga_survive(F.data_ptr(), k["ld"], k["n"], k["n_obj"], S, k["nsga2"], idx.data_ptr(),
                                  rank.data_ptr(), crowd.data_ptr(), ws.data_ptr(), stream()))
    assert np.array_equal(idx.cpu().numpy(), k["idx"])
    assert np.array_equal(rank.cpu().numpy(), k["rank"])
    assert np.array_equal(crowd.cpu().numpy(), k["crowd"])
    # gather of the survivors
    X = np.random.default_rng(1).normal(size=(k["n"], 7))
    Xo = torch.zeros(S, 7, dtype=torch.float64, device=DEV)
    Fo = torch.zeros(k["n_obj"], S + 2, dtype=torch.float32, device=DEV)
    check_ga(lib.glass_ga_gather(dev(X).data_ptr(), F.data_ptr(), k["ld"], idx.data_ptr(), S, 7, k["n_obj"],
                                 Xo.data_ptr(), Fo.data_ptr(), S + 2, stream()))
    assert np.array_equal(Xo.cpu().numpy(), X[k["idx"]])
    assert np.array_equal(Fo.cpu().numpy()[:, :S], k["F"][k["idx"]].T)


def test_bad_arguments_are_refused():
    lib = load_library()
    from clip_glass_b200._lib import GlassArgError
    t = torch.zeros(8, dtype=torch.float64, device=DEV)
    with pytest.raises(GlassArgError):
        check_ga(lib.glass_ga_permutations(t.data_ptr(), 5000, 1, t.data_ptr(), stream()))
    with pytest.raises(GlassArgError):
        check_ga(lib.glass_ga_survive(t.data_ptr(), 4, 8, 2, 4, 1, t.data_ptr(), t.data_ptr(), t.data_ptr(),
                                      t.data_ptr(), stream()))


def test_resident_generation_loop_on_a_known_problem():
    """The whole loop on the device with an analytic two-objective problem (Schaffer-like on the first variable,
    sphere on the rest): the population stays feasible and distinct, ranks / crowding agree with a host recomputation
    of the same F, and the front improves.  No host round trip inside ``step``."""
    from clip_glass_b200 import ga
    from clip_glass_b200.device_ga import DeviceGA
    P, V = 64, 32

    def evaluate(z32, f_cols, generation):
        rest = (z32[:, 1:] ** 2).sum(1)
        f_cols[0].copy_(z32[:, 0] ** 2 + rest)
        f_cols[1].copy_((z32[:, 0] - 2.0) ** 2 + rest)

    g = DeviceGA("nsga2", P, V, 2, -10.0, 10.0, evaluate, device=DEV, seed=5, pm_prob=None)
    X0 = np.random.default_rng(0).normal(0, 1, (P, V))
    g.initialize(X0)
    X, F, rank, crowd = g.population()
    F0 = np.stack([(X0.astype(np.float32) ** 2).sum(1),
                   ((X0[:, 0].astype(np.float32) - 2) ** 2 + (X0[:, 1:].astype(np.float32) ** 2).sum(1))], 1)
    idx, r_h, c_h = ga.rank_and_crowding_survival(F.astype(np.float64), P)
    assert sorted(map(tuple, X)) == sorted(map(tuple, X0))                  # a reordering of the sample
    np.testing.assert_allclose(np.sort(F[:, 0]), np.sort(F0[:, 0]), rtol=1e-5)
    assert np.array_equal(r_h, rank) and np.array_equal(idx, np.arange(P))  # already in survivor order
    assert np.array_equal(c_h, crowd)
    first = F.sum(1).mean()
    for _ in range(40):
        g.step()
    assert g.offspring_filled() == P
    X, F, rank, crowd = g.population()
    assert np.isfinite(X).all() and X.min() >= -10 and X.max() <= 10
    assert len({row.tobytes() for row in X}) == P                           # eliminate_duplicates
    Fh = np.stack([(X.astype(np.float32) ** 2).sum(1),
                   ((X[:, 0].astype(np.float32) - 2) ** 2 + (X[:, 1:].astype(np.float32) ** 2).sum(1))], 1)
    np.testing.assert_allclose(F, Fh, rtol=1e-4, atol=1e-5)                 # F travels with its row through survival
    assert F.sum(1).mean() < 0.5 * first
    # determinism: the same seed gives the same search, another seed a different one
    g2 = DeviceGA("nsga2", P, V, 2, -10.0, 10.0, evaluate, device=DEV, seed=5, pm_prob=None)
    g2.initialize(X0)
    for _ in range(40):
        g2.step()
    assert np.array_equal(g2.population()[0], X)
    g3 = DeviceGA("nsga2", P, V, 2, -10.0, 10.0, evaluate, device=DEV, seed=6, pm_prob=None)
    g3.initialize(X0)
    g3.step()
    g4 = DeviceGA("nsga2", P, V, 2, -10.0, 10.0, evaluate, device=DEV, seed=5, pm_prob=None)
    g4.initialize(X0)
    g4.step()
    assert not np.array_equal(g3.population()[0], g4.population()[0])


@pytest.mark.parametrize("config", ["StyleGAN2_ffhq_nod", "StyleGAN2_ffhq_d"])
def test_run_driver_with_device_resident_ga(tmp_path, config):
    """run.py's flow with --device-ga: the fitness engine scores the offspring straight from the buffers the GA
    kernels wrote (tiny network shapes, GA for _nod / NSGA-II for _d)."""
    from clip_glass_b200 import run as driver
    res = driver.main(["--device", DEV, "--config", config, "--generations", "4", "--save-each", "2",
                       "--tmp-folder", str(tmp_path), "--pop-size", "8", "--batch-size", "4", "--synthetic-seed", "100",
                       "--seed", "3", "--device-ga"], config_overrides=dict(gan_spec=W.TINY_GAN, clip_spec=W.TINY_CLIP))
    names = set(os.listdir(tmp_path))
    assert {"genetic-it-2.jpg", "genetic-it-final.jpg", "output.jpg", "genetic_result", "ls_result"} <= names, names
    with open(tmp_path / "genetic_result", "rb") as f:
        saved = pickle.load(f)
    F = np.atleast_2d(np.asarray(saved["F"], dtype=float))
    assert np.isfinite(F).all() and len(res.pop) == 8
    X = res.pop.get("X")
    assert X.shape == (8, 512) and np.abs(X).max() <= 10.0
