"""Genetic operators and the GA / NSGA-II generation loop behind ``operators.get_operators`` and ``run.py``.

The reference takes all of this from pymoo 0.4.2.1 (``run.py:6-9,59-76``, ``operators.py:5-7,45-59,69-70,75-77``),
which is neither vendored by the reference nor installable offline.  When pymoo imports, ``operators.get_operators``
and ``clip_glass_b200/run.py`` use it unchanged.  When it does not, the classes here stand in so that the driver
still runs end to end: they restate the published algorithms (Deb & Agrawal's simulated binary crossover and
polynomial mutation with pymoo's bounded variant, half-uniform crossover, bit-flip mutation, binary tournament,
fast non-dominated sort and crowding distance of NSGA-II) with pymoo's operator conventions **from recollection**:

    PARITY UNPINNED — there is no pymoo source on this box to check the random-number order or corner cases
    against.  tests/test_host_cpu.py pins the mathematical properties only (bounds, symmetry of SBX children around
    the parents' mean, distribution index behaviour, exact Pareto fronts on known sets).

Shapes follow pymoo: a crossover takes ``X[n_parents=2, n_matings, n_var]`` and returns ``[2, n_matings, n_var]``;
a mutation and a sampling work on ``[n, n_var]``.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional

import numpy as np


# ---------------------------------------------------------------------------
# operators
# ---------------------------------------------------------------------------
def _bounds(problem, n_var):
    xl = np.broadcast_to(np.asarray(problem.xl, dtype=float), (n_var,))
    xu = np.broadcast_to(np.asarray(problem.xu, dtype=float), (n_var,))
    return xl, xu


class SimulatedBinaryCrossover:
    """real_sbx (operators.py:69): bounded SBX, each variable recombined with probability 0.5."""

    def __init__(self, eta: float, prob: float = 0.9, prob_per_variable: float = 0.5, rng=None):
        self.eta, self.prob, self.prob_per_variable = float(eta), float(prob), prob_per_variable
        self.rng = rng or np.random
        self.n_parents = self.n_offsprings = 2

    def _betaq(self, beta, rand):
        alpha = 2.0 - np.power(beta, -(self.eta + 1.0))
        lo = rand <= 1.0 / alpha
        e = 1.0 / (self.eta + 1.0)
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.where(lo, np.power(rand * alpha, e), np.power(1.0 / (2.0 - rand * alpha), e))

    def _do(self, problem, X, **kw):
        X = X.astype(float)
        _, n_matings, n_var = X.shape
        xl, xu = _bounds(problem, n_var)
        do = self.rng.random((n_matings, n_var)) <= self.prob_per_variable
        do &= np.abs(X[0] - X[1]) > 1e-14
        y1, y2 = X.min(axis=0), X.max(axis=0)
        rand = self.rng.random((n_matings, n_var))
        delta = np.maximum(y2 - y1, 1e-10)
        c1 = 0.5 * ((y1 + y2) - self._betaq(1.0 + 2.0 * (y1 - xl) / delta, rand) * delta)
        c2 = 0.5 * ((y1 + y2) + self._betaq(1.0 + 2.0 * (xu - y2) / delta, rand) * delta)
        swap = self.rng.random((n_matings, n_var)) <= 0.5
        c1, c2 = np.where(swap, c2, c1), np.where(swap, c1, c2)
        out = X.copy()
        out[0][do] = c1[do]
        out[1][do] = c2[do]
        return np.clip(out, xl, xu)

    def do(self, problem, X, **kw):
        """Whole matings are recombined with probability ``prob`` (pymoo ``Crossover.do``)."""
        off = self._do(problem, X, **kw)
        keep = self.rng.random(X.shape[1]) >= self.prob
        off[:, keep] = X[:, keep]
        return off


class PolynomialMutation:
    """real_pm (operators.py:70): bounded polynomial mutation, per-variable probability ``prob``."""

    def __init__(self, eta: float, prob: Optional[float] = None, rng=None):
        self.eta, self.prob, self.rng = float(eta), prob, rng or np.random

    def _do(self, problem, X, **kw):
        X = X.astype(float)
        n, n_var = X.shape
        xl, xu = _bounds(problem, n_var)
        prob = self.prob if self.prob is not None else 1.0 / n_var
        do = self.rng.random(X.shape) < prob
        rand = self.rng.random(X.shape)
        span = xu - xl
        d1, d2 = (X - xl) / span, (xu - X) / span
        mp = 1.0 / (self.eta + 1.0)
        lo = rand <= 0.5
        v_lo = 2.0 * rand + (1.0 - 2.0 * rand) * np.power(1.0 - d1, self.eta + 1.0)
        v_hi = 2.0 * (1.0 - rand) + 2.0 * (rand - 0.5) * np.power(1.0 - d2, self.eta + 1.0)
        dq = np.where(lo, np.power(v_lo, mp) - 1.0, 1.0 - np.power(v_hi, mp))
        Y = np.where(do, np.clip(X + dq * span, xl, xu), X)
        return Y

    do = _do


class _IntegerFromFloat:
    """int_sbx / int_pm (operators.py:75-77): the real operator on [xl - 0.5, xu + 0.5) followed by rounding."""

    def __init__(self, inner):
        self.inner = inner
        self.n_parents = getattr(inner, "n_parents", None)

    class _Shift:
        def __init__(self, problem):
            self.xl = np.asarray(problem.xl, dtype=float) - (0.5 - 1e-16)
            self.xu = np.asarray(problem.xu, dtype=float) + (0.5 - 1e-16)

    def do(self, problem, X, **kw):
        Y = self.inner.do(self._Shift(problem), X.astype(float), **kw)
        xl, xu = _bounds(problem, Y.shape[-1])
        return np.clip(np.rint(Y), xl, xu).astype(int)

    _do = do


class IntegerRandomSampling:
    """int_random (operators.py:75): uniform integers in [xl, xu]."""

    def __init__(self, rng=None):
        self.rng = rng or np.random

    def _do(self, problem, n_samples, **kw):
        xl, xu = _bounds(problem, problem.n_var)
        return np.column_stack([self.rng.randint(int(xl[k]), int(xu[k]) + 1, size=n_samples)
                                for k in range(problem.n_var)])

    do = _do


class HalfUniformCrossover:
    """bin_hux (operators.py:52): exchange half of the differing genes."""

    def __init__(self, prob: float = 0.9, rng=None):
        self.prob, self.rng = float(prob), rng or np.random
        self.n_parents = self.n_offsprings = 2

    def _do(self, problem, X, **kw):
        _, n_matings, n_var = X.shape
        M = np.zeros((n_matings, n_var), dtype=bool)
        differ = X[0] != X[1]
        for i in range(n_matings):
            idx = np.where(differ[i])[0]
            n = math.ceil(len(idx) / 2)
            if n > 0:
                M[i, idx[self.rng.permutation(len(idx))[:n]]] = True
        out = X.copy()
        out[0][M], out[1][M] = X[1][M], X[0][M]
        return out

    def do(self, problem, X, **kw):
        off = self._do(problem, X, **kw)
        keep = self.rng.random(X.shape[1]) >= self.prob
        off[:, keep] = X[:, keep]
        return off


class BitflipMutation:
    """bin_bitflip (operators.py:58)."""

    def __init__(self, prob: Optional[float] = None, rng=None):
        self.prob, self.rng = prob, rng or np.random

    def _do(self, problem, X, **kw):
        prob = self.prob if self.prob is not None else 1.0 / X.shape[1]
        flip = self.rng.random(X.shape) < prob
        Y = X.astype(bool)
        return np.where(flip, ~Y, Y)

    do = _do


class _Sub:
    """A view of the problem restricted to some columns (mixed-variable wrappers)."""

    def __init__(self, problem, cols):
        self.n_var = len(cols)
        self.xl = np.broadcast_to(np.asarray(problem.xl, dtype=float), (problem.n_var,))[cols]
        self.xu = np.broadcast_to(np.asarray(problem.xu, dtype=float), (problem.n_var,))[cols]


class _Mixed:
    def __init__(self, mask: List[str], process: Dict[str, object]):
        self.mask = np.asarray(mask)
        self.process = process
        self.cols = {t: np.where(self.mask == t)[0] for t in process}


class MixedVariableSampling(_Mixed):       # operators.py:45-48
    def _do(self, problem, n_samples, **kw):
        X = np.empty((n_samples, len(self.mask)), dtype=object)
        for t, op in self.process.items():
            X[:, self.cols[t]] = op._do(_Sub(problem, self.cols[t]), n_samples, **kw)
        return X

    do = _do


class MixedVariableCrossover(_Mixed):      # operators.py:50-53
    n_parents = n_offsprings = 2

    def do(self, problem, X, **kw):
        out = np.empty(X.shape, dtype=object)
        for t, op in self.process.items():
            c = self.cols[t]
            sub = X[:, :, c].astype(bool if t == "bool" else float)
            out[:, :, c] = op.do(_Sub(problem, c), sub, **kw)
        return out


class MixedVariableMutation(_Mixed):       # operators.py:55-59
    def do(self, problem, X, **kw):
        out = np.empty(X.shape, dtype=object)
        for t, op in self.process.items():
            c = self.cols[t]
            out[:, c] = op.do(_Sub(problem, c), X[:, c].astype(bool if t == "bool" else float), **kw)
        return out

    _do = do


def get_crossover(name: str, **kw):
    if name == "real_sbx":
        return SimulatedBinaryCrossover(**kw)
    if name == "int_sbx":
        return _IntegerFromFloat(SimulatedBinaryCrossover(**kw))
    if name == "bin_hux":
        return HalfUniformCrossover(**kw)
    raise KeyError(name)


def get_mutation(name: str, **kw):
    if name == "real_pm":
        return PolynomialMutation(**kw)
    if name == "int_pm":
        return _IntegerFromFloat(PolynomialMutation(**kw))
    if name == "bin_bitflip":
        return BitflipMutation(**kw)
    raise KeyError(name)


def get_sampling(name: str, **kw):
    if name == "int_random":
        return IntegerRandomSampling(**kw)
    raise KeyError(name)


# ---------------------------------------------------------------------------
# survival: NSGA-II rank + crowding, GA fitness
# ---------------------------------------------------------------------------
def fast_non_dominated_sort(F: np.ndarray) -> List[np.ndarray]:
    """Fronts (lists of indices) of the minimisation problem F [n, n_obj] (Deb et al. 2002)."""
    n = F.shape[0]
    le = (F[:, None, :] <= F[None, :, :]).all(-1)
    lt = (F[:, None, :] < F[None, :, :]).any(-1)
    dom = le & lt                                   # dom[i, j]: i dominates j
    n_dom = dom.sum(0)
    fronts, cur = [], np.where(n_dom == 0)[0]
    assigned = np.zeros(n, dtype=bool)
    while len(cur):
        fronts.append(cur)
        assigned[cur] = True
        n_dom = n_dom - dom[cur].sum(0)
        cur = np.where((n_dom == 0) & ~assigned)[0]
    return fronts


def crowding_distance(F: np.ndarray) -> np.ndarray:
    n, m = F.shape
    if n <= 2:
        return np.full(n, np.inf)
    d = np.zeros(n)
    for k in range(m):
        order = np.argsort(F[:, k], kind="mergesort")
        f = F[order, k]
        span = f[-1] - f[0]
        d[order[0]] = d[order[-1]] = np.inf
        if span > 0:
            d[order[1:-1]] += (f[2:] - f[:-2]) / span
    return d


def rank_and_crowding_survival(F: np.ndarray, n_survive: int):
    """Indices of the survivors, their ranks and crowding distances (NSGA-II)."""
    survivors, ranks, crowd = [], [], []
    for r, front in enumerate(fast_non_dominated_sort(F)):
        cd = crowding_distance(F[front])
        if len(survivors) + len(front) > n_survive:
            order = np.argsort(-cd, kind="mergesort")[: n_survive - len(survivors)]
            front, cd = front[order], cd[order]
        survivors += list(front)
        ranks += [r] * len(front)
        crowd += list(cd)
        if len(survivors) >= n_survive:
            break
    return np.asarray(survivors), np.asarray(ranks), np.asarray(crowd)


class Individual:
    def __init__(self, X, F):
        self.X, self.F = X, F


class Population(list):
    def get(self, key):
        return np.stack([getattr(p, key) for p in self])


class Result:
    pass


class Algorithm:
    """The part of a pymoo algorithm object that run.py's callback reads: ``.pop`` (with ``.get('X')``; items with
    ``.X`` / ``.F``) and ``.problem``."""

    def __init__(self, name: str, pop_size: int, sampling, crossover, mutation, eliminate_duplicates=True,
                 callback: Optional[Callable] = None, seed: Optional[int] = None, **_):
        assert name in ("ga", "nsga2"), name
        self.name, self.pop_size = name, pop_size
        self.sampling, self.crossover, self.mutation = sampling, crossover, mutation
        self.eliminate_duplicates, self.callback = eliminate_duplicates, callback
        self.rng = np.random.RandomState(seed) if seed is not None else np.random
        self.pop, self.problem, self.n_gen = Population(), None, 0

    # -- pieces ------------------------------------------------------------
    def _evaluate(self, X):
        out = {}
        self.problem._evaluate(np.asarray(X, dtype=float) if X.dtype != object else X, out)
        F = np.asarray(out["F"], dtype=float)
        return F.reshape(len(X), -1)

    def _survive(self, X, F):
        if self.name == "nsga2":
            idx, rank, crowd = rank_and_crowding_survival(F, self.pop_size)
        else:
            idx = np.argsort(F[:, 0], kind="mergesort")[: self.pop_size]
            rank, crowd = np.zeros(len(idx), dtype=int), -F[idx, 0]
        self.pop = Population(Individual(X[i], F[i] if F.shape[1] > 1 else F[i, 0]) for i in idx)
        self._rank, self._crowd = rank, crowd

    def _tournament(self, n_select):
        """Binary tournament on (rank, crowding) for NSGA-II, on F for GA."""
        n = len(self.pop)
        P = np.concatenate([self.rng.permutation(n) for _ in range(math.ceil(2 * n_select / n))])[: 2 * n_select]
        a, b = P[0::2], P[1::2]
        better_a = (self._rank[a] < self._rank[b]) | ((self._rank[a] == self._rank[b]) & (self._crowd[a] >= self._crowd[b]))
        return np.where(better_a, a, b)

    def _mate(self, X, n_off, multiple=1):
        """Offspring that are not duplicates of the population (or of each other); the count is kept a multiple of
        ``multiple`` (the reference asserts pop % minibatch == 0, models.py:112).  A candidate is a duplicate when
        max |row - candidate| <= 1e-16 for a population row or an offspring accepted earlier; rows are screened on
        their first variable (a necessary condition) before the full comparison."""
        off = []
        Xf = np.asarray(X, dtype=float)
        acc = np.empty((n_off, X.shape[1]), dtype=float)          # float view of the accepted offspring

        def duplicate(cf, pool):
            near = np.flatnonzero(np.abs(pool[:, 0] - cf[0]) <= 1e-16)
            return near.size > 0 and bool((np.abs(pool[near] - cf).max(axis=1) <= 1e-16).any())

        for _ in range(100):
            need = n_off - len(off)
            if need <= 0:
                break
            n_matings = math.ceil(need / 2)
            parents = self._tournament(2 * n_matings).reshape(n_matings, 2)
            Xp = np.stack([X[parents[:, 0]], X[parents[:, 1]]])
            C = self.crossover.do(self.problem, Xp)
            C = C.reshape(-1, X.shape[1])
            C = self.mutation.do(self.problem, C)
            Cf = np.asarray(C, dtype=float)
            for c, cf in zip(C, Cf):
                if len(off) >= n_off:
                    break
                if self.eliminate_duplicates and (duplicate(cf, acc[:len(off)]) or duplicate(cf, Xf)):
                    continue
                acc[len(off)] = cf
                off.append(c)
        while len(off) % multiple:
            off.append(off[-1])
        return np.stack(off)

    # -- loop ----------------------------------------------------------------
    def solve(self, problem, n_gen: int, verbose: bool = False, offspring_multiple: int = 1):
        self.problem = problem
        X = np.asarray(self.sampling._do(problem, self.pop_size))
        F = self._evaluate(X)
        self._survive(X, F)
        self.n_gen = 1
        if self.callback:
            self.callback(self)
        while self.n_gen < n_gen:
            X, F = self.pop.get("X"), self.pop.get("F").reshape(len(self.pop), -1)
            off = self._mate(X, self.pop_size, offspring_multiple)
            Fo = self._evaluate(off)
            self._survive(np.concatenate([X, off]), np.concatenate([F, Fo]))
            self.n_gen += 1
            if verbose:
                print(f"{self.n_gen:5d} | best {np.min(self.pop.get('F'), axis=0)}")
            if self.callback:
                self.callback(self)
        res = Result()
        res.pop = self.pop
        F = self.pop.get("F").reshape(len(self.pop), -1)
        if F.shape[1] == 1:
            best = int(np.argmin(F[:, 0]))
            res.X, res.F = self.pop[best].X, F[best]
        else:
            front = fast_non_dominated_sort(F)[0]
            res.X, res.F = self.pop.get("X")[front], F[front]
        return res


def get_algorithm(name, **kw):
    return Algorithm(name, **kw)


def minimize(problem, algorithm, termination, save_history=False, verbose=False, seed=None, **kw):
    """run.py:70-76: ``minimize(problem, algorithm, ("n_gen", G), ...)``."""
    assert termination[0] == "n_gen"
    if seed is not None:
        algorithm.rng = np.random.RandomState(seed)
    batch = getattr(getattr(problem, "config", None), "batch_size", 1)
    return algorithm.solve(problem, int(termination[1]), verbose=verbose, offspring_multiple=batch)
