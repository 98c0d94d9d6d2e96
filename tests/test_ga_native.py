"""CPU checks of the genetic-operator arithmetic the CUDA kernels wrap (clip_glass_b200/csrc/ga_ops.cuh): the same
__host__ __device__ functions compiled with g++ (tests/native/ga_host.cpp) against the host operators of
clip_glass_b200/ga.py.  The kernels themselves are compared through the C ABI in tests/test_gpu_ga.py."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from tests import ga_cases as C

HERE = os.path.dirname(os.path.abspath(__file__))


class HostParams(ctypes.Structure):
    _fields_ = [("sbx_eta", ctypes.c_double), ("sbx_prob", ctypes.c_double), ("sbx_prob_var", ctypes.c_double),
                ("pm_eta", ctypes.c_double), ("pm_prob", ctypes.c_double), ("n_var", ctypes.c_int32),
                ("integer", ctypes.c_int32)]


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("ga_host") / "ga_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                    os.path.join(HERE, "native", "ga_host.cpp"), "-o", out], check=True)
    return ctypes.CDLL(out)


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_philox_known_answers(host):
    for ctr, key, want in C.PHILOX_KAT:
        c = np.asarray(ctr, dtype=np.uint32)
        host.ga_host_philox(ptr(c), ctypes.c_uint32(key[0]), ctypes.c_uint32(key[1]))
        assert tuple(int(v) for v in c) == want


def test_uniform_stream(host):
    n = 200001
    u = np.empty(n)
    host.ga_host_uniform(ctypes.c_uint64(7), ctypes.c_uint64(0), ptr(u), ctypes.c_int64(n))
    assert u.min() >= 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 4 * np.sqrt(1 / 12 / n) and abs(u.var() - 1 / 12) < 1e-3
    assert len(np.unique(u)) == n
    # the stream is addressed by (seed, pair index): a request split in two reproduces it
    tail = np.empty(n - 1000)
    host.ga_host_uniform(ctypes.c_uint64(7), ctypes.c_uint64(500), ptr(tail), ctypes.c_int64(n - 1000))
    assert np.array_equal(tail, u[1000:])
    other = np.empty(16)
    host.ga_host_uniform(ctypes.c_uint64(8), ctypes.c_uint64(0), ptr(other), ctypes.c_int64(16))
    assert not np.array_equal(other, u[:16])


def test_permutations_are_stable_argsorts(host):
    rng = np.random.default_rng(0)
    keys = rng.random((5, 64))
    keys[2, 10] = keys[2, 3]
    out = np.empty((5, 64), dtype=np.int32)
    host.ga_host_permutations(ptr(keys), 64, 5, ptr(out))
    assert np.array_equal(out, np.argsort(keys, axis=1, kind="stable"))
    assert all(sorted(row) == list(range(64)) for row in out)


@pytest.mark.parametrize("case", C.OFFSPRING_CASES, ids=lambda c: "int" if c["integer"] else f"real{c['V']}")
def test_offspring_against_host_operators(host, case):
    k = C.offspring_case(**case)
    p = HostParams(**k["params"])
    out = np.full_like(k["expect"], np.nan)
    host.ga_host_offspring(ctypes.byref(p), ptr(k["X"]), ptr(k["parents"]), ptr(k["bounds"]), ptr(k["rnd"]),
                           k["M"], ptr(out))
    assert out.shape == (2 * k["M"], case["V"])
    if case["integer"]:
        assert np.array_equal(out, k["expect"])
    else:
        # same libm on both sides: identical up to the vectorised pow numpy may use
        np.testing.assert_allclose(out, k["expect"], rtol=1e-14, atol=1e-14)
        changed = np.abs(out - np.concatenate([k["X"][k["parents"][:, 0]], k["X"][k["parents"][:, 1]]])) > 0
        assert 0.3 < changed.mean() < 1.0
    assert out.min() >= case["xl"] and out.max() <= case["xu"]


def test_tournament_against_host(host):
    for seed, n, n_select in ((0, 64, 64), (1, 10, 24), (2, 512, 512)):
        k = C.tournament_case(seed, n, n_select)
        sel = np.empty(n_select, dtype=np.int32)
        host.ga_host_tournament(ptr(k["pairs"]), ptr(k["rank"]), ptr(k["crowd"]), n_select, ptr(sel))
        assert np.array_equal(sel, k["expect"])


@pytest.mark.parametrize("have,n_off", [(0, 16), (3, 8), (2, 64)])
def test_duplicate_elimination_against_host_loop(host, have, n_off):
    k = C.dedup_case(seed=5, n_x=12, n_c=14, n_off=n_off, have=have, V=24)
    off = k["off"].copy()
    z32 = np.zeros(off.shape, dtype=np.float32)
    n_have = np.asarray([have], dtype=np.int32)
    flags, dest = np.empty(14, dtype=np.int32), np.empty(14, dtype=np.int32)
    host.ga_host_dedup_append(ptr(k["cand"]), 14, ptr(k["X"]), 12, ptr(off), n_off, ptr(n_have), 24,
                              ctypes.c_double(1e-16), 1, ptr(z32), ptr(flags), ptr(dest))
    assert int(n_have[0]) == len(k["expect"])
    assert np.array_equal(off[: n_have[0]], k["expect"])
    assert np.array_equal(z32[have: n_have[0]], k["expect"][have:].astype(np.float32))
    assert flags[1] == 1 and flags[3] == 1 and flags[5] == 0 and (flags[4] == 1) == (have > 0)


@pytest.mark.parametrize("case", C.SURVIVE_CASES, ids=lambda c: f"n{c['n']}m{c['n_obj']}s{c['n_survive']}")
def test_survival_against_host(host, case):
    k = C.survive_case(**case)
    S = k["n_survive"]
    idx, rank, crowd = np.full(S, -1, dtype=np.int32), np.full(S, -1, dtype=np.int32), np.full(S, np.nan)
    host.ga_host_survive(ptr(k["Fcm"]), k["ld"], k["n"], k["n_obj"], S, k["nsga2"], ptr(idx), ptr(rank), ptr(crowd))
    assert np.array_equal(idx, k["idx"])
    assert np.array_equal(rank, k["rank"])
    assert np.array_equal(crowd, k["crowd"])            # same IEEE operations in the same order: exact, incl. inf


class _HostLibShim:
    """The glass_ga_* entry points served by the host compilation of the same arithmetic, on CPU tensors: lets the
    CPU suite drive clip_glass_b200.device_ga.DeviceGA's buffer handling (the product class itself has no host path:
    the test swaps the library and the CUDA queries underneath it)."""

    def __init__(self, host):
        self.h = host
        self.h.ga_host_rand_count.restype = ctypes.c_int64

    @staticmethod
    def _p(v):
        return ctypes.c_void_p(v) if isinstance(v, int) else v

    def glass_ga_rand_count(self, M, V):
        return self.h.ga_host_rand_count(M, V)

    def glass_ga_dedup_workspace(self, n):
        return n * 8

    def glass_ga_survive_workspace(self, n):
        return 16

    def glass_ga_uniform(self, seed, offset, out, n, stream):
        self.h.ga_host_uniform(ctypes.c_uint64(seed), ctypes.c_uint64(offset), self._p(out), ctypes.c_int64(n))
        return 0

    def glass_ga_permutations(self, keys, n, n_perm, out, stream):
        self.h.ga_host_permutations(self._p(keys), n, n_perm, self._p(out))
        return 0

    def glass_ga_tournament(self, pairs, rank, crowd, n_select, sel, stream):
        self.h.ga_host_tournament(self._p(pairs), self._p(rank), self._p(crowd), n_select, self._p(sel))
        return 0

    def glass_ga_offspring(self, params, X, parents, bounds, rnd, M, out, stream):
        g = params._obj
        hp = HostParams(g.sbx_eta, g.sbx_prob, g.sbx_prob_var, g.pm_eta, g.pm_prob, g.n_var, g.integer)
        self.h.ga_host_offspring(ctypes.byref(hp), self._p(X), self._p(parents), self._p(bounds), self._p(rnd), M,
                                 self._p(out))
        return 0

    def glass_ga_dedup_append(self, cand, n_c, X, n_x, off, n_off, n_have, V, eps, eliminate, z32, ws, stream):
        flags = np.empty(n_c, dtype=np.int32)
        dest = np.empty(n_c, dtype=np.int32)
        self.h.ga_host_dedup_append(self._p(cand), n_c, self._p(X), n_x, self._p(off), n_off, self._p(n_have), V,
                                    ctypes.c_double(eps), eliminate, self._p(z32), ptr(flags), ptr(dest))
        return 0

    def glass_ga_pad(self, off, n_off, n_have, V, z32, stream):
        self.h.ga_host_pad(self._p(off), n_off, self._p(n_have), V, self._p(z32))
        return 0

    def glass_ga_cast_f32(self, x, z, n, stream):
        self.h.ga_host_cast(self._p(x), self._p(z), ctypes.c_int64(n))
        return 0

    def glass_ga_survive(self, F, ld, n, n_obj, S, nsga2, idx, rank, crowd, ws, stream):
        self.h.ga_host_survive(self._p(F), ld, n, n_obj, S, nsga2, self._p(idx), self._p(rank), self._p(crowd))
        return 0

    def glass_ga_gather(self, X, F, ld_in, idx, n_out, V, n_obj, Xo, Fo, ld_out, stream):
        self.h.ga_host_gather(self._p(X), self._p(F), ld_in, self._p(idx), n_out, V, n_obj, self._p(Xo), self._p(Fo),
                              ld_out)
        return 0


def _schaffer(z32, f_cols, generation):
    rest = (z32[:, 1:] ** 2).sum(1)
    f_cols[0].copy_(z32[:, 0] ** 2 + rest)
    f_cols[1].copy_((z32[:, 0] - 2.0) ** 2 + rest)


def test_resident_loop_bookkeeping_on_host_shim(host, monkeypatch):
    """DeviceGA's generation loop (buffer ping-pong, offsets, draw counter, survivor bookkeeping) with the kernels
    replaced by their host compilation."""
    import torch
    from clip_glass_b200 import device_ga as D, ga
    monkeypatch.setattr(D, "load_library", lambda: _HostLibShim(host))
    monkeypatch.setattr(D.torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(D.DeviceGA, "_stream", lambda self: None)
    P, V = 64, 32
    X0 = np.random.default_rng(0).normal(0, 1, (P, V))

    def run(seed, gens, algorithm="nsga2", n_obj=2, fn=_schaffer):
        g = D.DeviceGA(algorithm, P, V, n_obj, -10.0, 10.0, fn, device="cpu", seed=seed, pm_prob=None)
        g.initialize(X0)
        first = tuple(np.array(a) for a in g.population())       # (CPU tensors: .numpy() is a view)
        for _ in range(gens):
            g.step()
            assert g.offspring_filled() == P
        return g, first

    g, (X1, F1, rank1, crowd1) = run(5, 40)
    assert sorted(map(tuple, X1)) == sorted(map(tuple, X0))
    idx, r_h, c_h = ga.rank_and_crowding_survival(F1, P)
    assert np.array_equal(idx, np.arange(P)) and np.array_equal(r_h, rank1) and np.array_equal(c_h, crowd1)
    X, F, rank, crowd = g.population()
    assert np.isfinite(X).all() and X.min() >= -10 and X.max() <= 10
    assert len({row.tobytes() for row in X}) == P
    z = torch.from_numpy(X).float()
    cols = [torch.empty(P), torch.empty(P)]
    _schaffer(z, cols, 0)
    assert np.array_equal(F, torch.stack(cols, 1).numpy().astype(np.float64))     # F travels with its row
    idx, r_h, c_h = ga.rank_and_crowding_survival(F, P)
    whole = rank < rank.max()            # the last front was cut by crowding: its distances are those of the full front
    assert np.array_equal(idx, np.arange(P)) and np.array_equal(r_h, rank) and np.array_equal(c_h[whole], crowd[whole])
    assert F.sum(1).mean() < 0.5 * F1.sum(1).mean()
    assert np.array_equal(run(5, 40)[0].population()[0], X)                         # deterministic in the seed
    assert not np.array_equal(run(6, 1)[0].population()[0], run(5, 1)[0].population()[0])
    # single-objective GA (config 1): the population is sorted by F and improves
    def sphere(z32, f_cols, generation):
        f_cols[0].copy_((z32 ** 2).sum(1))
    g1, (_, Fa, _, _) = run(2, 30, "ga", 1, sphere)
    Fb = g1.population()[1][:, 0]
    assert np.all(np.diff(Fb) >= 0) and Fb.mean() < 0.5 * Fa[:, 0].mean()


def test_device_algorithm_follows_run_py_callback_contract(host, monkeypatch):
    """DeviceAlgorithm.solve as run.py uses it: the callback is called once per generation and finds ``.pop`` current
    on the saving generations; the result carries the Pareto front (NSGA-II) or the best individual (GA)."""
    from types import SimpleNamespace
    from clip_glass_b200 import device_ga as D, ga
    monkeypatch.setattr(D, "load_library", lambda: _HostLibShim(host))
    monkeypatch.setattr(D.torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(D.DeviceGA, "_stream", lambda self: None)
    monkeypatch.setattr(D.torch, "device", lambda *a: "cpu")

    import torch
    P, V = 16, 12
    for name, n_obj in (("nsga2", 2), ("ga", 1)):
        def fn(z, outs, gen, first_group):
            cols = [torch.empty(len(z)) for _ in range(2)]
            _schaffer(z, cols, gen)
            for o, c in zip(outs, cols):         # the GA config keeps the first objective only
                o.copy_(c)
        monkeypatch.setattr(D, "engine_evaluator", lambda e, b, s=0: D.sharded_evaluator(fn, b))
        cfg = SimpleNamespace(problem_args=dict(n_obj=n_obj), use_discriminator=n_obj == 2, batch_size=4, noise_seed=0)
        problem = SimpleNamespace(config=cfg, n_var=V, xl=np.full(V, -10.0), xu=np.full(V, 10.0),
                                  generator=SimpleNamespace(engine=SimpleNamespace(device=0), _last_rows={1: 2}))
        problem.generator.forget_population = lambda g=problem.generator: setattr(g, "_last_rows", None)
        sampling = SimpleNamespace(_do=lambda prob, n: np.random.default_rng(1).normal(0, 1, (n, prob.n_var)))
        seen = []

        def cb(alg):
            seen.append((alg.n_gen, None if not len(alg.pop) else alg.pop.get("F").copy()))
        alg = D.DeviceAlgorithm(name, pop_size=P, sampling=sampling, callback=cb, callback_each=3, seed=4, pm_prob=None)
        res = alg.solve(problem, 7)
        assert [g for g, _ in seen] == list(range(1, 8))
        assert problem.generator._last_rows is None
        X, F, _, _ = alg.state.population()
        assert np.array_equal(alg.pop.get("X"), X) and len(res.pop) == P
        Fpop = alg.pop.get("F").reshape(P, -1)
        assert np.array_equal(Fpop, F)
        fresh = {g: f for g, f in seen if g in (3, 6, 7)}
        assert all(f is not None for f in fresh.values()) and not np.array_equal(fresh[3], fresh[6])
        if n_obj == 2:
            front = ga.fast_non_dominated_sort(F)[0]
            assert np.array_equal(np.atleast_2d(res.X), X[front]) and np.array_equal(res.F, F[front])
        else:
            assert np.array_equal(res.X, X[0]) and res.F[0] == F[:, 0].min()


def _ga_gloo_worker(rank, world, port, so_path, q):
    import torch
    import torch.distributed as tdist
    from clip_glass_b200 import device_ga as D
    tdist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    host = ctypes.CDLL(so_path)
    D.load_library = lambda: _HostLibShim(host)
    D.torch.cuda.is_available = lambda: True
    D.DeviceGA._stream = lambda self: None
    calls = []

    def local(z, outs, gen, first_group):
        calls.append((len(z), first_group))
        _schaffer(z, outs, gen)
    P, V = 12, 8                                  # three minibatches of four over two ranks: shards of 8 and 4
    g = D.DeviceGA("nsga2", P, V, 2, -10.0, 10.0, D.sharded_evaluator(local, 4), device="cpu", seed=9, pm_prob=None)
    g.initialize(np.random.default_rng(2).normal(0, 1, (P, V)))
    for _ in range(5):
        g.step()
    X, F, rank_, crowd = g.population()
    q.put((rank, np.array(X), np.array(F), calls))
    tdist.destroy_process_group()


def test_resident_loop_sharded_over_two_ranks_gloo(host, tmp_path):
    """World size 2 over gloo: every rank runs the same seeded search, evaluates only its shard of the offspring and
    one all-gather rebuilds F — the populations agree with each other and with a single-process run."""
    import torch.multiprocessing as mp
    so = os.path.join(str(tmp_path), "ga_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                    os.path.join(HERE, "native", "ga_host.cpp"), "-o", so], check=True)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ga_gloo_worker, args=(r, 2, 29893, so, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    (_, X0, F0, calls0), (_, X1, F1, calls1) = res
    assert np.array_equal(X0, X1) and np.array_equal(F0, F1)
    assert calls0 == [(8, 0)] * 6 and calls1 == [(4, 2)] * 6
    # single process, same seed
    from clip_glass_b200 import device_ga as D
    import unittest.mock as um
    with um.patch.object(D, "load_library", lambda: _HostLibShim(host)), \
            um.patch.object(D.torch.cuda, "is_available", lambda: True), \
            um.patch.object(D.DeviceGA, "_stream", lambda self: None):
        g = D.DeviceGA("nsga2", 12, 8, 2, -10.0, 10.0,
                       D.sharded_evaluator(lambda z, o, gen, fg: _schaffer(z, o, gen), 4), device="cpu", seed=9,
                       pm_prob=None)
        g.initialize(np.random.default_rng(2).normal(0, 1, (12, 8)))
        for _ in range(5):
            g.step()
        assert np.array_equal(g.population()[0], X0)


def _mate_reference(alg, X, n_off, multiple=1):
    """The plain loop ga.Algorithm._mate replaces (every accepted offspring and every population row compared in
    full): the screened version must accept exactly the same candidates from the same random draws."""
    import math
    off = []
    for _ in range(100):
        need = n_off - len(off)
        if need <= 0:
            break
        n_matings = math.ceil(need / 2)
        parents = alg._tournament(2 * n_matings).reshape(n_matings, 2)
        Xp = np.stack([X[parents[:, 0]], X[parents[:, 1]]])
        C = alg.mutation.do(alg.problem, alg.crossover.do(alg.problem, Xp).reshape(-1, X.shape[1]))
        for c in C:
            if len(off) >= n_off:
                break
            cf = np.asarray(c, dtype=float)
            if alg.eliminate_duplicates and (
                    any(np.abs(np.asarray(o, dtype=float) - cf).max() <= 1e-16 for o in off)
                    or (np.abs(np.asarray(X, dtype=float) - cf).max(axis=1) <= 1e-16).any()):
                continue
            off.append(c)
    while len(off) % multiple:
        off.append(off[-1])
    return np.stack(off)


@pytest.mark.parametrize("kind", ["real", "real_dups", "int", "mixed"])
def test_host_mating_duplicate_screen_equals_plain_loop(kind):
    from types import SimpleNamespace
    from clip_glass_b200 import ga

    def build(seed):
        rs = np.random.RandomState(seed)
        if kind == "int":             # GPT-2 operators on a tiny vocabulary: equal first genes and duplicates are common
            V, xl, xu = 3, 0.0, 2.0
            cross = ga._IntegerFromFloat(ga.SimulatedBinaryCrossover(3.0, 1.0, rng=rs))
            mut = ga._IntegerFromFloat(ga.PolynomialMutation(3.0, 0.5, rng=rs))
            X = rs.randint(0, 3, size=(12, V))
        elif kind == "mixed":         # BigGAN's mixed real / bool population (object rows)
            V = 10
            mask = ["real"] * 4 + ["bool"] * 6
            cross = ga.MixedVariableCrossover(mask, {"real": ga.SimulatedBinaryCrossover(3.0, 0.3, rng=rs),
                                                     "bool": ga.HalfUniformCrossover(0.2, rng=rs)})
            mut = ga.MixedVariableMutation(mask, {"real": ga.PolynomialMutation(3.0, 0.05, rng=rs),
                                                  "bool": ga.BitflipMutation(0.01, rng=rs)})
            X = np.empty((12, V), dtype=object)
            X[:, :4] = rs.normal(size=(12, 4))
            X[:, 4:] = rs.random_sample((12, 6)) < 0.3
            xl, xu = -2.0, 2.0
        else:                         # StyleGAN2 operators; "real_dups": matings mostly kept, mutation rare => duplicates
            V, xl, xu = 16, -10.0, 10.0
            cross = ga.SimulatedBinaryCrossover(3.0, 1.0 if kind == "real" else 0.2, rng=rs)
            mut = ga.PolynomialMutation(3.0, 0.5 if kind == "real" else 0.01, rng=rs)
            X = rs.normal(size=(12, V))
        alg = ga.Algorithm("nsga2", pop_size=12, sampling=None, crossover=cross, mutation=mut)
        alg.rng = rs
        alg.problem = SimpleNamespace(n_var=V, xl=np.full(V, xl), xu=np.full(V, xu))
        alg.pop = [None] * 12
        alg._rank, alg._crowd = rs.randint(0, 3, size=12), rs.random_sample(12)
        return alg, X

    rejected = 0
    for seed in range(6):
        a, X = build(seed)
        b, Xb = build(seed)
        got = a._mate(X, 12, 4)
        want = _mate_reference(b, Xb, 12, 4)
        assert got.shape == want.shape and np.array_equal(np.asarray(got, dtype=float), np.asarray(want, dtype=float))
        # both consumed the same number of draws
        assert a.rng.random_sample() == b.rng.random_sample()
        Xf, gf = np.asarray(X, dtype=float), np.asarray(got, dtype=float)
        if kind != "real":
            c, _ = build(seed)
            c.eliminate_duplicates = False
            rejected += int(not np.array_equal(np.asarray(c._mate(X, 12, 4), dtype=float), gf))
        assert not (np.abs(gf[:, None, :] - Xf[None]).max(-1) <= 1e-16).any()      # no offspring repeats a parent row
    assert kind == "real" or rejected > 0, "the scenario never produced a duplicate"


def test_run_driver_device_ga_branch_on_host_shim(host, monkeypatch, tmp_path):
    """run.py's --device-ga branch (DeviceAlgorithm in place of pymoo's minimize) end to end on the CPU: the kernels
    are served by their host compilation, the fitness engine by an analytic function."""
    import pickle
    import torch
    from types import SimpleNamespace
    from clip_glass_b200 import device_ga as D, run as driver
    from tests.test_host_cpu import _StubProblem, _DRIVER_ARGS

    class Problem(_StubProblem):
        def __init__(self, config):
            super().__init__(config)
            self.generator.engine = SimpleNamespace(device=0)
            self.generator.forgot = 0
            self.generator.forget_population = lambda: setattr(self.generator, "forgot", self.generator.forgot + 1)

    def fitness(z, outs, gen, first_group):
        outs[0].copy_((z ** 2).sum(1))
        if len(outs) > 1:
            outs[1].copy_((z[:, 0] - 1).abs())

    monkeypatch.setattr(driver, "GenerationProblem", Problem)
    monkeypatch.setattr(D, "load_library", lambda: _HostLibShim(host))
    monkeypatch.setattr(D.torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(D.DeviceGA, "_stream", lambda self: None)
    monkeypatch.setattr(D.torch, "device", lambda *a: "cpu")
    monkeypatch.setattr(D, "engine_evaluator", lambda e, b, s=0: D.sharded_evaluator(fitness, b))
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    for config_name in ("StyleGAN2_ffhq_d", "StyleGAN2_ffhq_nod"):
        folder = tmp_path / config_name
        res = driver.main(_DRIVER_ARGS + ["--config", config_name, "--tmp-folder", str(folder), "--seed", "3",
                                          "--device-ga"])
        names = set(os.listdir(folder))
        assert {"genetic-it-2.jpg", "genetic-it-4.jpg", "genetic-it-final.jpg", "output.jpg", "genetic_result",
                "ls_result"} <= names, names
        with open(folder / "genetic_result", "rb") as f:
            saved = pickle.load(f)
        assert np.isfinite(np.asarray(saved["F"], dtype=float)).all() and len(res.pop) == 8
        X = res.pop.get("X")
        assert X.shape == (8, 512) and np.abs(X).max() <= 10.0
