"""Seeded cases for the genetic operators, with the expected results computed by the host operators of
clip_glass_b200/ga.py.  Shared by tests/test_ga_native.py (the operator arithmetic of csrc/ga_ops.cuh compiled for the
host) and tests/test_gpu_ga.py (the CUDA kernels through the C ABI)."""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np

from clip_glass_b200 import ga

SHIFT = 0.5 - 1e-16          # ga._IntegerFromFloat._Shift


class ReplayRng:
    """Hands out pre-drawn uniforms in the order the host operators ask for them."""

    def __init__(self, flat: np.ndarray):
        self.flat, self.at = np.asarray(flat, dtype=np.float64), 0

    def random(self, shape=None):
        shape = (shape,) if isinstance(shape, int) else tuple(shape)
        n = int(np.prod(shape))
        out = self.flat[self.at:self.at + n].reshape(shape)
        self.at += n
        return out


def rand_count(M: int, V: int) -> int:
    return 7 * M * V + M


def offspring_case(seed: int, n: int, M: int, V: int, integer: bool, sbx_prob: float, pm_prob, xl: float, xu: float,
                   uniforms: np.ndarray = None):
    """Population, parent indices, uniforms and the children ga.py's crossover + mutation make from them."""
    rng = np.random.default_rng(seed)
    if integer:
        X = rng.integers(int(xl), int(xu) + 1, size=(n, V)).astype(np.float64)
    else:
        X = np.clip(rng.normal(0.0, 1.0, size=(n, V)) * 3.0, xl, xu)
        X[1] = X[0]                                   # identical parents: |x0 - x1| <= 1e-14 keeps the genes
        X[2, : V // 2] = xl                           # genes on the bounds
        X[3, : V // 2] = xu
    parents = rng.integers(0, n, size=(M, 2)).astype(np.int32)
    parents[0] = (0, 1)
    parents[1 % M] = (2, 3)
    rnd = rng.random(rand_count(M, V)) if uniforms is None else np.asarray(uniforms, dtype=np.float64)
    assert rnd.size == rand_count(M, V)
    problem = SimpleNamespace(n_var=V, xl=np.full(V, xl), xu=np.full(V, xu))
    replay = ReplayRng(rnd)
    sbx = ga.SimulatedBinaryCrossover(eta=3.0, prob=sbx_prob, rng=replay)
    pm = ga.PolynomialMutation(eta=3.0, prob=pm_prob, rng=replay)
    cross, mut = (ga._IntegerFromFloat(sbx), ga._IntegerFromFloat(pm)) if integer else (sbx, pm)
    Xp = np.stack([X[parents[:, 0]], X[parents[:, 1]]])
    C = cross.do(problem, Xp).reshape(-1, V)
    expect = np.asarray(mut.do(problem, C), dtype=np.float64)
    assert replay.at == rnd.size
    if integer:
        bounds = np.stack([np.full(V, xl - SHIFT), np.full(V, xu + SHIFT), np.full(V, xl), np.full(V, xu)])
    else:
        bounds = np.stack([np.full(V, xl), np.full(V, xu), np.full(V, xl), np.full(V, xu)])
    params = dict(sbx_eta=3.0, sbx_prob=sbx_prob, sbx_prob_var=0.5, pm_eta=3.0,
                  pm_prob=-1.0 if pm_prob is None else pm_prob, n_var=V, integer=int(integer))
    return dict(X=X, parents=parents, rnd=rnd, bounds=np.ascontiguousarray(bounds), params=params, expect=expect, M=M)


OFFSPRING_CASES = [
    # the reference's StyleGAN2 operators (operators.py:66-71; config.py:91-92 bounds +-10)
    dict(seed=1, n=16, M=8, V=512, integer=False, sbx_prob=1.0, pm_prob=0.5, xl=-10.0, xu=10.0),
    # pymoo defaults: some matings kept, mutation probability 1 / n_var
    dict(seed=2, n=12, M=9, V=33, integer=False, sbx_prob=0.9, pm_prob=None, xl=-2.0, xu=2.0),
    # the GPT-2 integer operators (operators.py:73-78; tokens 0..50256)
    dict(seed=3, n=16, M=8, V=20, integer=True, sbx_prob=1.0, pm_prob=0.5, xl=0.0, xu=50256.0),
]


def tournament_case(seed: int, n: int, n_select: int):
    rng = np.random.default_rng(seed)
    rank = rng.integers(0, 3, size=n).astype(np.int32)
    crowd = rng.random(n)
    crowd[rng.integers(0, n, size=max(1, n // 4))] = np.inf
    crowd[1] = crowd[0]                                                    # tie: the first of the pair wins
    perms = [rng.permutation(n) for _ in range(math.ceil(2 * n_select / n))]

    class _Rng:
        def __init__(self):
            self.k = 0

        def permutation(self, m):
            assert m == n
            self.k += 1
            return perms[self.k - 1]

    alg = ga.Algorithm("nsga2", pop_size=n, sampling=None, crossover=None, mutation=None)
    alg.rng, alg.pop, alg._rank, alg._crowd = _Rng(), [None] * n, rank, crowd
    expect = alg._tournament(n_select)
    pairs = np.concatenate(perms)[: 2 * n_select].astype(np.int32)
    return dict(pairs=pairs, rank=rank, crowd=crowd, expect=expect.astype(np.int32), n_select=n_select)


def dedup_case(seed: int, n_x: int, n_c: int, n_off: int, have: int, V: int):
    """Candidates with planted duplicates; expected buffer from a direct restatement of ga.Algorithm._mate's loop."""
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n_x, V))
    off = np.zeros((n_off, V))
    off[:have] = rng.normal(size=(have, V))
    cand = rng.normal(size=(n_c, V))
    cand[1] = X[n_x - 1]                       # equals a population row
    cand[3] = cand[0]                          # equals an earlier candidate
    if have:
        cand[4] = off[have - 1]                # equals an accepted offspring
    cand[5] = X[0]
    cand[5, V - 1] += 1e-9                     # differs in the last variable only: kept
    acc = [o for o in off[:have]]
    for c in cand:
        if len(acc) >= n_off:
            break
        if any(np.abs(o - c).max() <= 1e-16 for o in acc) or (np.abs(X - c).max(axis=1) <= 1e-16).any():
            continue
        acc.append(c)
    expect = np.stack(acc)
    return dict(X=X, off=off, cand=cand, have=have, n_off=n_off, expect=expect)


def survive_case(seed: int, n: int, n_obj: int, n_survive: int, nsga2: bool, quantize: int = 0):
    """Random objectives (optionally quantised so that ties and duplicate points occur) and ga.py's survivors."""
    rng = np.random.default_rng(seed)
    F = rng.normal(size=(n, n_obj)).astype(np.float32)
    if quantize:
        F = (np.round(F * quantize) / quantize).astype(np.float32)
    if nsga2:
        idx, rank, crowd = ga.rank_and_crowding_survival(F.astype(np.float64), n_survive)
    else:
        idx = np.argsort(F[:, 0].astype(np.float64), kind="mergesort")[:n_survive]
        rank, crowd = np.zeros(n_survive, dtype=int), -F[idx, 0].astype(np.float64)
    ld = n + 3                                  # a leading dimension larger than n, as the driver uses
    Fcm = np.zeros((n_obj, ld), dtype=np.float32)
    Fcm[:, :n] = F.T
    return dict(F=F, Fcm=Fcm, ld=ld, n=n, n_obj=n_obj, n_survive=n_survive, nsga2=int(nsga2),
                idx=np.asarray(idx, dtype=np.int32), rank=np.asarray(rank, dtype=np.int32),
                crowd=np.asarray(crowd, dtype=np.float64))


SURVIVE_CASES = [
    dict(seed=11, n=128, n_obj=2, n_survive=64, nsga2=True),               # config 2: P = 64 merged with 64 offspring
    dict(seed=12, n=64, n_obj=2, n_survive=64, nsga2=True),                # first generation: everything survives
    dict(seed=13, n=200, n_obj=2, n_survive=77, nsga2=True, quantize=4),   # ties, duplicate points, many fronts
    dict(seed=14, n=96, n_obj=3, n_survive=40, nsga2=True, quantize=8),
    dict(seed=15, n=1024, n_obj=2, n_survive=512, nsga2=True),             # config 4: P = 512
    dict(seed=16, n=16, n_obj=1, n_survive=8, nsga2=False),                # config 1: GA, n_obj = 1
    dict(seed=17, n=130, n_obj=1, n_survive=64, nsga2=False, quantize=4),
    dict(seed=18, n=5, n_obj=2, n_survive=3, nsga2=True),                  # fronts of one or two points
]

# Philox4x32-10 known-answer vectors (Random123 kat_vectors: counter, key, output)
PHILOX_KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]
