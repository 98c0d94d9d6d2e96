"""Mirror of the reference's problem.py: the pymoo ``Problem`` plugin.

``GenerationProblem._evaluate(x, out)`` has the reference's contract
(problem.py:14-29): ``x`` float64 [P, n_var] from pymoo; sets ``out["F"]``
([P] = -sim for n_obj 1, [P,2] = (-sim, hinge) for NSGA-II) and
``out["G"]`` = zeros[P].  Two equivalent routes:

  fused=True  (default)  one ``glass_evaluate_host`` call per generation
                         (H2D of x, all kernels, D2H of F);
  fused=False            the reference's three façade calls
                         (generate -> clip_similarity -> discriminate).

With ``torch.distributed`` initialised the population is sharded over ranks
(clip_glass_b200/dist.py) and F is all-gathered, so every rank returns the
full F exactly as a single-GPU run would.
"""
from __future__ import annotations

import numpy as np
import torch

from .generator import Generator
from .utils import nvtx_range

try:                                    # pymoo==0.4.2.1 is not installable offline
    from pymoo.model.problem import Problem as _Base
    HAVE_PYMOO = True
except Exception:                       # pragma: no cover - depends on the box
    HAVE_PYMOO = False

    class _Base:                        # the attributes pymoo's Problem.__init__ sets and run.py reads
        def __init__(self, n_var=-1, n_obj=-1, n_constr=0, xl=None, xu=None, **kwargs):
            self.n_var, self.n_obj, self.n_constr = n_var, n_obj, n_constr
            self.xl = np.full(n_var, xl, dtype=float) if np.isscalar(xl) else xl
            self.xu = np.full(n_var, xu, dtype=float) if np.isscalar(xu) else xu

        def evaluate(self, x, *args, **kwargs):
            out = {}
            self._evaluate(np.atleast_2d(x), out, *args, **kwargs)
            return out


class GenerationProblem(_Base):
    def __init__(self, config, generator=None):
        self.generator = generator if generator is not None else Generator(config)
        self.config = config
        self.generation = 0
        super().__init__(**self.config.problem_args)

    def _two_objective(self):
        return self.config.problem_args["n_obj"] == 2 and self.config.use_discriminator

    def _evaluate(self, x, out, *args, **kwargs):
        with nvtx_range("glass._evaluate"):
            return self._evaluate_generation(x, out, *args, **kwargs)

    def _evaluate_generation(self, x, out, *args, **kwargs):
        self.generation += 1
        if self.config.task == "img2txt":
            # problem.py:15-29 for the GPT-2 config: generate -> clip_similarity, F = -sim.  With
            # config.clip_token_map == "standin" (no vocabulary files on the box) the text round trip is replaced by
            # the documented token-level stand-in (models.standin_clip_tokens); the GPU work is identical.
            from . import dist

            def local_text(xs, first_group):
                ls = self.config.latent(self.config)
                ls.set_from_population(xs)
                if getattr(self.config, "clip_token_map", None) == "standin":
                    from .models import standin_clip_tokens
                    gen = self.generator.model.parse_out_tokens(self.generator.model.generate_tokens(ls()[0]))
                    sim = self.generator.clip_similarity(standin_clip_tokens(gen, self.generator.text_spec))
                else:
                    generated = self.generator.generate(ls, minibatch=self.config.batch_size)
                    sim = self.generator.clip_similarity(generated)
                return -sim.cpu().numpy().astype(np.float32), None

            # candidates are independent on this path (the reference decodes the whole population in one batch and
            # ignores `minibatch`, models.py:46): shards of any size, one all-gather of F
            neg_sim, _ = dist.sharded_evaluate(x, 1, 1, local_text)
            out["F"] = neg_sim
            out["G"] = np.zeros((x.shape[0]))
            return
        if getattr(self.config, "fused", True):
            from . import dist
            seed = int(getattr(self.config, "noise_seed", 0)) + self.generation
            self.generator.engine.set_batch_size(self.config.batch_size)
            noise = kwargs.get("noise")     # explicit noise: [groups][layers] tensors (tests / parity runs)

            multi = dist._world()[1] > 1

            def local(xs, first_group):
                nz = None
                if noise is not None:
                    nz = noise[first_group:first_group + xs.shape[0] // self.config.batch_size]
                self.generator.remember_population(xs)      # the engine keeps this shard's images (save_callback)
                if multi:
                    # outputs stay on the device: the all-gather reads them there (one D2H of the gathered F)
                    z = torch.from_numpy(np.ascontiguousarray(xs, dtype=np.float64)).to(
                        self.config.device, non_blocking=True).float()           # latent.py:38
                    return self.generator.engine.evaluate_device(z, noise=nz, seed=seed, first_group=first_group)
                return self.generator.engine.evaluate(xs, noise=nz, seed=seed, first_group=first_group)

            n_obj = 2 if self._two_objective() else 1
            neg_sim, hinge = dist.sharded_evaluate(x, self.config.batch_size, n_obj, local)
            if self._two_objective():
                out["F"] = np.column_stack((neg_sim, hinge))
            else:
                out["F"] = neg_sim
            out["G"] = np.zeros((x.shape[0]))
            return
        # the reference's own call sequence (problem.py:15-29)
        ls = self.config.latent(self.config)
        ls.set_from_population(x)
        with torch.no_grad():
            generated = self.generator.generate(ls, minibatch=self.config.batch_size)
            sim = self.generator.clip_similarity(generated).cpu().numpy()
            if self._two_objective():
                dis = self.generator.discriminate(generated, minibatch=self.config.batch_size)
                hinge = torch.relu(1 - dis)
                hinge = hinge.squeeze(1).cpu().numpy()
                out["F"] = np.column_stack((-sim, hinge))
            else:
                out["F"] = -sim
            out["G"] = np.zeros((x.shape[0]))
