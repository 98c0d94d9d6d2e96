"""Mirror of the reference's driver (run.py:15-125) on top of the B200 fitness path.

    python -m clip_glass_b200.run --config StyleGAN2_ffhq_d --target "<text>" [--generations 500] ...

Same six flags (run.py:17-22), same flow: config -> GenerationProblem -> get_operators -> get_algorithm ->
minimize(("n_gen", G)) with ``save_callback`` every ``--save-each`` generations -> ``genetic_result`` pickle,
``ls_result`` state dict, decision making on the Pareto front, ``output.jpg``.  pymoo 0.4.2.1 drives the search when
it is importable (the reference's own engine); otherwise ``clip_glass_b200.ga`` stands in (parity unpinned, see its
header).  Extra flags of this build: ``--pop-size`` / ``--batch-size`` (the reference has no CLI flag for them;
BASELINE.json's populations need them), ``--synthetic-seed`` (seeded random weights: there are no checkpoints
offline), ``--seed``, ``--device-ga`` (the GA / NSGA-II operators run on the GPU and the population stays in
device memory between generations, clip_glass_b200/device_ga.py; StyleGAN2 configs).  Under ``torchrun`` the driver
creates the process group itself (NCCL, collective timeout: ``dist.init_from_env``), every rank runs the same seeded
search, the population is sharded inside ``_evaluate`` (clip_glass_b200/dist.py) and rank 0 writes the files.
"""
from __future__ import annotations

import argparse
import os
import pickle

import numpy as np
import torch

from .config import get_config
from .operators import HAVE_PYMOO, get_operators
from .problem import GenerationProblem

if HAVE_PYMOO:                                           # pragma: no cover - pymoo is absent offline
    from pymoo.factory import get_algorithm, get_decision_making, get_decomposition
    from pymoo.optimize import minimize
else:
    from .ga import get_algorithm, minimize
    get_decision_making = get_decomposition = None


def pseudo_weights_choice(F: np.ndarray, weights=(0, 1)) -> int:
    """run.py:107-112 picks one Pareto point with pymoo's pseudo-weights decision making, falling back to ASF
    decomposition.  Without pymoo: the point whose pseudo-weight vector (normalised distance to the worst value of
    each objective) is closest to ``weights``."""
    F = np.atleast_2d(np.asarray(F, dtype=float))
    span = F.max(0) - F.min(0)
    span[span == 0] = 1.0
    pw = (F.max(0) - F) / span
    s = pw.sum(1, keepdims=True)
    s[s == 0] = 1.0
    pw = pw / s
    return int(np.argmin(np.abs(pw - np.asarray(weights, dtype=float)).sum(1)))


def main(argv=None, config_overrides=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--device", type=str, default="cuda")
    parser.add_argument("--config", type=str, default="StyleGAN2_ffhq_d")
    parser.add_argument("--generations", type=int, default=500)
    parser.add_argument("--save-each", type=int, default=50)
    parser.add_argument("--tmp-folder", type=str, default="./tmp")
    parser.add_argument("--target", type=str, default="a wolf at night with the moon in the background")
    parser.add_argument("--pop-size", type=int, default=None)
    parser.add_argument("--batch-size", type=int, default=None)
    parser.add_argument("--synthetic-seed", type=int, default=None)
    parser.add_argument("--seed", type=int, default=None)
    parser.add_argument("--device-ga", action="store_true")
    config = parser.parse_args(argv)
    vars(config).update(get_config(config.config))                         # run.py:25
    if config.pop_size is None or config.batch_size is None:
        raise SystemExit("config lacks pop_size / batch_size")
    for k in ("pop_size", "batch_size", "synthetic_seed"):
        v = getattr(parser.parse_args(argv), k)
        if v is not None:
            setattr(config, k, v)
    vars(config).update(config_overrides or {})
    if getattr(config, "synthetic_seed", None) is not None and getattr(config, "text_features", None) is None \
            and config.task == "txt2img":
        # no checkpoints offline: the cached CLIP.encode_text(target) (generator.py:23-24) is a seeded stand-in
        config.text_features = torch.randn(1, 512, generator=torch.Generator().manual_seed(config.synthetic_seed + 5))
    # torchrun: one rank per GPU, every rank runs the same seeded search, the population is sharded inside _evaluate
    from . import dist
    rank, world, local_rank = dist.init_from_env(config.device, timeout_s=getattr(config, "collective_timeout", 600.0))
    if world > 1:
        if str(config.device).startswith("cuda"):
            config.device = "cuda:%d" % local_rank
        if config.seed is None:
            config.seed = 0                       # the ranks must draw the same populations
        if getattr(config, "max_population", None) is None:
            # each rank's workspace holds its own shard (Generator renders a whole population in chunks when saving)
            config.max_population = max(b - a for a, b in dist.shard_bounds(config.pop_size, config.batch_size, world))
    if config.seed is not None:
        np.random.seed(config.seed)
        torch.manual_seed(config.seed)

    state = dict(iteration=0)

    def save_callback(algorithm):                                          # run.py:29-51
        state["iteration"] += 1
        it = state["iteration"]
        if (it % config.save_each == 0 or it == config.generations) and rank == 0:      # one rank writes the files
            if config.problem_args["n_obj"] == 1:
                X = np.stack([p.X for p in sorted(algorithm.pop, key=lambda p: p.F)])
            else:
                X = algorithm.pop.get("X")
            ls = config.latent(config)
            ls.set_from_population(X)
            with torch.no_grad():
                generated = algorithm.problem.generator.generate(ls, minibatch=config.batch_size)
                ext = "jpg" if config.task == "txt2img" else "txt"
                name = "genetic-it-%d.%s" % (it, ext) if it < config.generations else "genetic-it-final.%s" % ext
                algorithm.problem.generator.save(generated, os.path.join(config.tmp_folder, name))

    problem = GenerationProblem(config)                                    # run.py:54
    operators = get_operators(config)                                      # run.py:55
    os.makedirs(config.tmp_folder, exist_ok=True)
    if config.device_ga:
        # same operators and parameters (operators.py:66-71), run by the glass_ga_* kernels on the resident population
        if config.config.split("_")[0] != "StyleGAN2":
            raise SystemExit("--device-ga drives the StyleGAN2 configs (real-valued latents)")
        from .device_ga import DeviceAlgorithm
        algorithm = DeviceAlgorithm(config.algorithm, pop_size=config.pop_size, sampling=operators["sampling"],
                                    callback=save_callback, callback_each=config.save_each,
                                    seed=config.seed or 0, eliminate_duplicates=True,
                                    sbx_eta=3.0, sbx_prob=1.0, pm_eta=3.0, pm_prob=0.5)
        res = algorithm.solve(problem, config.generations, verbose=True)
    else:
        algorithm = get_algorithm(config.algorithm, pop_size=config.pop_size, sampling=operators["sampling"],
                                  crossover=operators["crossover"], mutation=operators["mutation"],
                                  eliminate_duplicates=True, callback=save_callback)
        res = minimize(problem, algorithm, ("n_gen", config.generations), save_history=False, verbose=True,
                       **({"seed": config.seed} if config.seed is not None else {}))
    if rank != 0:
        return res
    with open(os.path.join(config.tmp_folder, "genetic_result"), "wb") as f:          # run.py:79-84
        pickle.dump(dict(X=res.X, F=res.F, G=getattr(res, "G", None), CV=getattr(res, "CV", None)), f)
    if config.problem_args["n_obj"] == 1:                                  # run.py:92-96
        X = np.stack([p.X for p in sorted(res.pop, key=lambda p: p.F)])
    else:
        X = res.pop.get("X")
    ls = config.latent(config)
    ls.set_from_population(X)
    torch.save(ls.state_dict() if hasattr(ls, "state_dict") else {}, os.path.join(config.tmp_folder, "ls_result"))
    if config.problem_args["n_obj"] == 1:                                  # run.py:103-115
        X = np.atleast_2d(res.X)
    else:
        if get_decision_making is not None:                                # pragma: no cover
            try:
                result = get_decision_making("pseudo-weights", [0, 1]).do(res.F)
            except Exception:
                print("Warning: cant use pseudo-weights")
                result = get_decomposition("asf").do(res.F, [0, 1]).argmin()
        else:
            result = pseudo_weights_choice(res.F, (0, 1))
        X = np.atleast_2d(np.atleast_2d(res.X)[result])
    ls.set_from_population(X)
    with torch.no_grad():
        generated = problem.generator.generate(ls)                         # run.py:117-118
    ext = "jpg" if config.task == "txt2img" else "txt"
    problem.generator.save(generated, os.path.join(config.tmp_folder, "output.%s" % ext))
    return res


if __name__ == "__main__":
    main()
