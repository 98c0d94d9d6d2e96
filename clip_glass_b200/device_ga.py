"""GPU-resident GA / NSGA-II generation loop.

The reference's loop (run.py:59-76: pymoo ``get_algorithm`` + ``minimize``) keeps the population on the host and
sends all of it through ``GenerationProblem._evaluate`` every generation (latent.py:38 H2D of the latents,
problem.py:20,24 D2H of the fitnesses).  Here the population ``X`` (f64 [P, n_var]), its objectives, ranks and
crowding distances live in device memory for the whole search; one generation is

    uniforms -> permutations -> binary tournament -> SBX + polynomial mutation -> duplicate elimination
    -> fitness (``glass_evaluate_device`` on the fp32 copy the elimination kernel writes)
    -> rank-and-crowding (or single-objective) survival -> gather of the survivors

as ``glass_ga_*`` launches on the current stream (clip_glass_b200/csrc/ga.cu) with no host synchronisation inside;
the host reads ``F`` only when a callback or the final result asks for it.  Operator arithmetic and conventions are
those of ``clip_glass_b200/ga.py`` (pymoo 0.4.2.1 itself is absent offline: parity with it is unpinned, see that
file); tests/test_gpu_ga.py checks every kernel against ga.py on the same uniform draws.  Differences from the host
loop, by construction: the random numbers come from a Philox4x32-10 stream addressed by (seed, draw index) instead of
numpy's Mersenne twister; every mating round makes a full set of ``pop_size`` candidates (the host loop shrinks the
round to the number still missing) and at most ``rounds`` rounds run, after which missing rows repeat the last
accepted offspring.  The first population is sampled by the reference's own host operator (operators.py:17-25) and
uploaded once.

There is no host fallback: without the CUDA library and a GPU the constructor raises.
"""
from __future__ import annotations

import ctypes
import math
from typing import Callable, List, Optional

import numpy as np
import torch

from ._lib import GlassGaParams, GlassNoise, check, check_ga, load_library
from .utils import nvtx_range


class DeviceGA:
    """Population state + one-generation step on the device.

    ``evaluate(z32, f_cols, generation)`` must enqueue, on the current stream, the fitness of the fp32 candidates
    ``z32`` [pop_size, n_var] into the ``n_obj`` device vectors ``f_cols`` (fp32 [pop_size] each).
    """

    def __init__(self, algorithm: str, pop_size: int, n_var: int, n_obj: int, xl, xu, evaluate: Callable,
                 device="cuda", seed: int = 0, integer: bool = False, sbx_eta: float = 3.0, sbx_prob: float = 1.0,
                 sbx_prob_var: float = 0.5, pm_eta: float = 3.0, pm_prob: Optional[float] = 0.5,
                 eliminate_duplicates: bool = True, rounds: int = 3):
        assert algorithm in ("ga", "nsga2"), algorithm
        assert algorithm == "nsga2" or n_obj == 1, "the single-objective GA ranks by F[:, 0]"
        if not torch.cuda.is_available():
            raise RuntimeError("DeviceGA needs a CUDA device (there is no host fallback; use clip_glass_b200.ga)")
        if not 0 < 2 * int(pop_size) <= 4096:
            raise ValueError("DeviceGA handles populations of 1..2048 (the merged 2 * pop_size candidates are ranked "
                             "by one thread block, glass_ga_survive)")
        if not 0 < int(n_obj) <= min(8, int(n_var)):
            raise ValueError("n_obj must be in 1..min(8, n_var)")
        self.lib = load_library()
        self.device = torch.device(device)
        self.algorithm, self.P, self.V, self.n_obj = algorithm, int(pop_size), int(n_var), int(n_obj)
        self.evaluate, self.seed, self.rounds = evaluate, int(seed), int(rounds)
        self.eliminate = bool(eliminate_duplicates)
        self.M = math.ceil(self.P / 2)                       # matings per round; 2M candidates
        self.n_perm = math.ceil(4 * self.M / self.P)
        self.params = GlassGaParams(sbx_eta, sbx_prob, sbx_prob_var, pm_eta, -1.0 if pm_prob is None else pm_prob,
                                    self.V, int(integer))
        xl = np.broadcast_to(np.asarray(xl, dtype=np.float64), (self.V,))
        xu = np.broadcast_to(np.asarray(xu, dtype=np.float64), (self.V,))
        shift = (0.5 - 1e-16) if integer else 0.0            # ga._IntegerFromFloat._Shift
        d, P, V, M = self.device, self.P, self.V, self.M
        self.bounds = torch.from_numpy(np.stack([xl - shift, xu + shift, xl, xu])).to(d)
        f64, f32, i32 = torch.float64, torch.float32, torch.int32
        self.X = [torch.zeros(2 * P, V, dtype=f64, device=d) for _ in range(2)]      # [population ; offspring]
        self.F = [torch.zeros(n_obj, 2 * P, dtype=f32, device=d) for _ in range(2)]  # column-major objectives
        self.cur = 0
        self.z32 = torch.zeros(P, V, dtype=f32, device=d)
        self.rank = torch.zeros(P, dtype=i32, device=d)
        self.crowd = torch.zeros(P, dtype=f64, device=d)
        self.idx = torch.zeros(P, dtype=i32, device=d)
        self.keys = torch.zeros(self.n_perm * P, dtype=f64, device=d)
        self.perms = torch.zeros(self.n_perm * P, dtype=i32, device=d)
        self.sel = torch.zeros(2 * M, dtype=i32, device=d)
        self.n_rand = int(self.lib.glass_ga_rand_count(M, V))
        self.rnd = torch.zeros(self.n_rand, dtype=f64, device=d)
        self.cand = torch.zeros(2 * M, V, dtype=f64, device=d)
        self.n_have = torch.zeros(1, dtype=i32, device=d)
        self.ws_dedup = torch.zeros(int(self.lib.glass_ga_dedup_workspace(2 * M)), dtype=torch.uint8, device=d)
        self.ws_survive = torch.zeros(int(self.lib.glass_ga_survive_workspace(2 * P)), dtype=torch.uint8, device=d)
        self.generation = 0
        self.launches = 0
        self._draws = 0                                       # Philox pair counter: every draw of the search is new

    # -- pieces ------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _uniform(self, out: torch.Tensor) -> None:
        check_ga(self.lib.glass_ga_uniform(self.seed, self._draws, out.data_ptr(), out.numel(), self._stream()))
        self._draws += (out.numel() + 1) // 2
        self.launches += 1

    def _survive(self, n: int) -> None:
        """Survivors of rows [0, n) of the current buffers -> rows [0, P) of the other pair; rank / crowd updated."""
        X, F = self.X[self.cur], self.F[self.cur]
        Xn, Fn = self.X[1 - self.cur], self.F[1 - self.cur]
        s = self._stream()
        check_ga(self.lib.glass_ga_survive(F.data_ptr(), 2 * self.P, n, self.n_obj, self.P,
                                           int(self.algorithm == "nsga2"), self.idx.data_ptr(), self.rank.data_ptr(),
                                           self.crowd.data_ptr(), self.ws_survive.data_ptr(), s))
        check_ga(self.lib.glass_ga_gather(X.data_ptr(), F.data_ptr(), 2 * self.P, self.idx.data_ptr(), self.P, self.V,
                                          self.n_obj, Xn.data_ptr(), Fn.data_ptr(), 2 * self.P, s))
        self.launches += 2
        self.cur = 1 - self.cur

    def _f_cols(self, first_row: int) -> List[torch.Tensor]:
        F = self.F[self.cur]
        return [F[k, first_row:first_row + self.P] for k in range(self.n_obj)]

    # -- the loop ------------------------------------------------------------
    def initialize(self, X0: np.ndarray) -> None:
        """Generation 1: the sampled population (host, operators.py:9-34) is uploaded, evaluated and ranked."""
        X0 = np.ascontiguousarray(X0, dtype=np.float64)
        assert X0.shape == (self.P, self.V), X0.shape
        self.X[self.cur][: self.P].copy_(torch.from_numpy(X0), non_blocking=False)
        check_ga(self.lib.glass_ga_cast_f32(self.X[self.cur].data_ptr(), self.z32.data_ptr(), self.P * self.V,
                                            self._stream()))
        self.launches += 1
        self.generation = 1
        self.evaluate(self.z32, self._f_cols(0), self.generation)
        self._survive(self.P)

    def mate(self) -> None:
        """Offspring rows [P, 2P) of the current X buffer (and their fp32 copy in ``z32``)."""
        P, V, M = self.P, self.V, self.M
        X = self.X[self.cur]
        off_ptr = X.data_ptr() + P * V * 8
        self.n_have.zero_()
        for _ in range(self.rounds):
            s = self._stream()
            self._uniform(self.keys)
            check_ga(self.lib.glass_ga_permutations(self.keys.data_ptr(), P, self.n_perm, self.perms.data_ptr(), s))
            check_ga(self.lib.glass_ga_tournament(self.perms.data_ptr(), self.rank.data_ptr(), self.crowd.data_ptr(),
                                                  2 * M, self.sel.data_ptr(), s))
            self._uniform(self.rnd)
            check_ga(self.lib.glass_ga_offspring(ctypes.byref(self.params), X.data_ptr(), self.sel.data_ptr(),
                                                 self.bounds.data_ptr(), self.rnd.data_ptr(), M, self.cand.data_ptr(),
                                                 s))
            check_ga(self.lib.glass_ga_dedup_append(self.cand.data_ptr(), 2 * M, X.data_ptr(), P, off_ptr, P,
                                                    self.n_have.data_ptr(), V, 1e-16, int(self.eliminate),
                                                    self.z32.data_ptr(), self.ws_dedup.data_ptr(), s))
            self.launches += 3 + (3 if self.eliminate else 2)
        check_ga(self.lib.glass_ga_pad(off_ptr, P, self.n_have.data_ptr(), V, self.z32.data_ptr(), self._stream()))
        self.launches += 1

    def step(self) -> None:
        """One generation; everything is enqueued on the current stream, nothing is read back."""
        assert self.generation >= 1, "call initialize() first"
        with nvtx_range("glass.ga.mate"):
            self.mate()
        self.generation += 1
        with nvtx_range("glass.ga.fitness"):
            self.evaluate(self.z32, self._f_cols(self.P), self.generation)
        with nvtx_range("glass.ga.survive"):
            self._survive(2 * self.P)

    # -- read-back (synchronises) ---------------------------------------------
    def population(self):
        """(X f64 [P, n_var], F f64 [P, n_obj], rank, crowd) of the current population, on the host."""
        X = self.X[self.cur][: self.P].cpu().numpy()
        F = self.F[self.cur][:, : self.P].t().contiguous().cpu().numpy().astype(np.float64)
        return X, F, self.rank.cpu().numpy(), self.crowd.cpu().numpy()

    def offspring_filled(self) -> int:
        """Distinct offspring the last ``mate()`` accepted before padding (synchronises; diagnostics)."""
        return int(self.n_have.item())


def sharded_evaluator(local: Callable, batch_size: int):
    """``evaluate`` callback that cuts the offspring into per-rank shards (whole minibatches, dist.shard_bounds), calls
    ``local(z32_shard, f_cols_shard, generation, first_group)`` for this rank's shard and rebuilds every rank's
    objectives with ONE all-gather (dist.py) — the same sharding ``GenerationProblem._evaluate`` uses, on the buffers
    the GA kernels read.  Single process: no collective."""
    from . import dist

    def evaluate(z32: torch.Tensor, f_cols, generation: int) -> None:
        rank, world = dist._world()
        pop = z32.shape[0]
        bounds = dist.shard_bounds(pop, batch_size, world)
        lo, hi = bounds[rank]
        if hi > lo:
            local(z32[lo:hi], [c[lo:hi] for c in f_cols], generation, lo // batch_size)
        if world > 1:
            longest = max(b - a for a, b in bounds)
            send = torch.zeros(len(f_cols), longest, dtype=torch.float32, device=z32.device)
            for k, c in enumerate(f_cols):
                send[k, : hi - lo] = c[lo:hi]
            recv = dist._gather_blocks(send, world)
            for r, (a, b) in enumerate(bounds):
                for k, c in enumerate(f_cols):
                    c[a:b] = recv[r, k, : b - a]

    return evaluate


def engine_evaluator(engine, batch_size: int, noise_seed: int = 0):
    """The fused fitness path (``glass_evaluate_device``) on the GA's device buffers: fresh noise per generation
    (seed + generation, indexed by the global minibatch group, as problem.GenerationProblem._evaluate draws it)."""
    lib = load_library()

    def local(z32: torch.Tensor, outs, generation: int, first_group: int) -> None:
        engine.set_batch_size(batch_size)
        nz = GlassNoise()
        nz.seed = (int(noise_seed) + generation) & 0xFFFFFFFFFFFFFFFF
        nz.first_group, nz.noise, nz.noise_on_device = first_group, None, 0
        stream = ctypes.c_void_p(torch.cuda.current_stream(z32.device).cuda_stream)
        hinge = outs[1].data_ptr() if len(outs) > 1 else None
        check(lib.glass_evaluate_device(engine._h, z32.data_ptr(), z32.shape[0], ctypes.byref(nz), outs[0].data_ptr(),
                                        hinge, stream))

    return sharded_evaluator(local, batch_size)


class _Individual:
    def __init__(self, X, F):
        self.X, self.F = X, F


class _Population(list):
    def get(self, key):
        return np.stack([getattr(p, key) for p in self])


class DeviceAlgorithm:
    """What run.py's callback and result handling read of a pymoo algorithm (``.pop``, ``.problem``), driven by
    ``DeviceGA``.  ``solve`` mirrors ``ga.Algorithm.solve`` / pymoo's ``minimize(problem, algorithm, ("n_gen", G))``;
    the population is copied to the host only for callbacks that read it (``callback_each``) and for the result."""

    def __init__(self, name: str, pop_size: int, sampling, callback=None, callback_each: int = 1, seed: int = 0,
                 eliminate_duplicates: bool = True, integer: bool = False, **operator_kw):
        self.name, self.pop_size, self.sampling = name, pop_size, sampling
        self.callback, self.callback_each, self.seed = callback, max(1, int(callback_each)), seed
        self.eliminate_duplicates, self.integer, self.operator_kw = eliminate_duplicates, integer, operator_kw
        self.pop, self.problem, self.n_gen, self.state = _Population(), None, 0, None

    def _sync_pop(self):
        X, F, _, _ = self.state.population()
        self.pop = _Population(_Individual(X[i], F[i] if F.shape[1] > 1 else F[i, 0]) for i in range(len(X)))

    def solve(self, problem, n_gen: int, verbose: bool = False):
        from . import ga
        self.problem = problem
        cfg = problem.config
        n_obj = 2 if (cfg.problem_args["n_obj"] == 2 and cfg.use_discriminator) else 1
        engine = problem.generator.engine
        self.state = DeviceGA(self.name, self.pop_size, problem.n_var, n_obj, problem.xl, problem.xu,
                              engine_evaluator(engine, cfg.batch_size, int(getattr(cfg, "noise_seed", 0))),
                              device=torch.device("cuda", engine.device), seed=self.seed, integer=self.integer,
                              eliminate_duplicates=self.eliminate_duplicates, **self.operator_kw)
        self.state.initialize(np.asarray(self.sampling._do(problem, self.pop_size), dtype=np.float64))
        self.n_gen = 1
        # the engine's cached images are those of the offspring it scored last, not of rows the host has seen
        problem.generator.forget_population()
        while True:
            if self.callback:
                # run.py's save_callback counts its calls and reads .pop only on saving generations
                if self.n_gen % self.callback_each == 0 or self.n_gen == n_gen:
                    self._sync_pop()
                    if verbose:
                        print(f"{self.n_gen:5d} | best {np.min(self.pop.get('F').reshape(len(self.pop), -1), axis=0)}")
                self.callback(self)
            if self.n_gen >= n_gen:
                break
            self.state.step()
            self.n_gen += 1
        self._sync_pop()
        res = ga.Result()
        res.pop = self.pop
        F = self.pop.get("F").reshape(len(self.pop), -1)
        if F.shape[1] == 1:
            best = int(np.argmin(F[:, 0]))
            res.X, res.F = self.pop[best].X, F[best]
        else:
            front = ga.fast_non_dominated_sort(F)[0]
            res.X, res.F = self.pop.get("X")[front], F[front]
        return res
