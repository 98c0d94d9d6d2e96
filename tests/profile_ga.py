"""Measurement of the search loop at the benchmarked shape (StyleGAN2_ffhq_d, P = 64, batch 4, synthetic weights):
generations per second with the host GA (clip_glass_b200/ga.py operators + GenerationProblem._evaluate with host
buffers: the reference's data flow, run.py:59-76) and with the GPU-resident GA (clip_glass_b200/device_ga.py: the
population never leaves the device).  Prints one JSON line; not a bench.py line (bench.py measures the fitness step).

    python tests/profile_ga.py [--pop 64] [--gens 20]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pop", type=int, default=64)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--gens", type=int, default=20)
    ap.add_argument("--tiny", action="store_true", help="reduced network shapes (plumbing check)")
    args = ap.parse_args()
    from clip_glass_b200 import ga, weights as W
    from clip_glass_b200.config import make_namespace
    from clip_glass_b200.device_ga import DeviceGA, engine_evaluator
    from clip_glass_b200.operators import get_operators
    from clip_glass_b200.problem import GenerationProblem

    P = args.pop
    text = torch.randn(1, 512, generator=torch.Generator().manual_seed(5))
    extra = dict(gan_spec=W.TINY_GAN, clip_spec=W.TINY_CLIP) if args.tiny else {}
    config = make_namespace("StyleGAN2_ffhq_d", device="cuda:0", target="synthetic", pop_size=P, batch_size=args.batch,
                            max_population=P, synthetic_seed=1000, text_features=text, noise_seed=7, **extra)
    problem = GenerationProblem(config)
    eng = problem.generator.engine
    ops = get_operators(config)
    np.random.seed(1)
    X0 = np.asarray(ops["sampling"]._do(problem, P), dtype=np.float64)

    # host GA: the reference's flow (operators on the host, the whole population through _evaluate every generation)
    alg = ga.Algorithm("nsga2", pop_size=P, sampling=ops["sampling"], crossover=ops["crossover"],
                       mutation=ops["mutation"], eliminate_duplicates=True, seed=1)
    alg.problem = problem
    alg._survive(X0, alg._evaluate(X0))
    for phase in ("warm", "timed"):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        t_ops = 0.0
        for _ in range(3 if phase == "warm" else args.gens):
            X, F = alg.pop.get("X"), alg.pop.get("F").reshape(P, -1)
            a = time.perf_counter()
            off = alg._mate(X, P, args.batch)
            t_ops += time.perf_counter() - a
            Fo = alg._evaluate(off)
            a = time.perf_counter()
            alg._survive(np.concatenate([X, off]), np.concatenate([F, Fo]))
            t_ops += time.perf_counter() - a
        torch.cuda.synchronize()
        host_ms = (time.perf_counter() - t0) * 1e3 / max(1, args.gens)
        host_ops_ms = t_ops * 1e3 / max(1, args.gens)

    # resident GA
    g = DeviceGA("nsga2", P, problem.n_var, 2, problem.xl, problem.xu,
                 engine_evaluator(eng, args.batch, 7), device="cuda:0", seed=1)
    g.initialize(X0)
    for _ in range(3):
        g.step()
    torch.cuda.synchronize()
    l0, e0 = g.launches, eng.launch_count
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    start.record()
    for _ in range(args.gens):
        g.step()
    stop.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.gens
    dev_ms = start.elapsed_time(stop) / args.gens
    filled = g.offspring_filled()
    X, F, rank, crowd = g.population()
    # the GA kernels alone (no fitness): same launches with a no-op evaluation
    g2 = DeviceGA("nsga2", P, problem.n_var, 2, problem.xl, problem.xu, lambda z, f, gen: None, device="cuda:0", seed=1)
    g2.X[0][:P].copy_(torch.from_numpy(X))
    g2.F[0][:, :P].copy_(torch.from_numpy(F.T.astype(np.float32)))
    g2.F[0][:, P:].copy_(torch.from_numpy(F.T.astype(np.float32)))
    g2.generation = 1
    g2._survive(P)
    for _ in range(3):
        g2.step()
    torch.cuda.synchronize()
    start.record()
    for _ in range(args.gens):
        g2.step()
    stop.record()
    torch.cuda.synchronize()
    ga_only_ms = start.elapsed_time(stop) / args.gens
    print(json.dumps({
        "workload": "StyleGAN2_ffhq_d NSGA-II generation" + (" (tiny nets)" if args.tiny else ""), "pop": P,
        "batch": args.batch, "generations_timed": args.gens,
        "host_ga_ms_per_generation": round(host_ms, 3), "host_ga_operator_ms": round(host_ops_ms, 3),
        "resident_ga_ms_per_generation": round(dev_ms, 3), "resident_ga_wall_ms_per_generation": round(wall_ms, 3),
        "resident_ga_kernels_only_ms": round(ga_only_ms, 4),
        "ga_launches_per_generation": (g.launches - l0) // args.gens,
        "engine_launches_per_generation": (eng.launch_count - e0) // args.gens,
        "offspring_filled": filled, "finite": bool(np.isfinite(F).all() and np.isfinite(X).all()),
        "h2d_bytes_per_generation": {"host_ga": P * problem.n_var * 8, "resident_ga": 0},
        "d2h_bytes_per_generation": {"host_ga": P * 2 * 4, "resident_ga": 0},
        "best_neg_sim": float(F[:, 0].min()), "front_size": int((rank == 0).sum()),
    }))


if __name__ == "__main__":
    main()
