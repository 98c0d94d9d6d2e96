"""Mirror of the reference's operators.py.

The three ``Sampling`` subclasses (operators.py:9-34) are pure numpy/scipy and
are restated here (with ``float``/``bool`` for the ``np.float``/``np.bool``
aliases that numpy>=1.24 removed).  SBX / PM / HUX / bit-flip live in pymoo
0.4.2.1, which is not installable offline: ``get_operators`` wires them exactly
as operators.py:37-82 from pymoo's factories when pymoo is importable, and from
the restatements in ``clip_glass_b200/ga.py`` otherwise (published algorithms,
pymoo conventions from recollection — parity unpinned).
"""
from __future__ import annotations

import numpy as np

try:
    from pymoo.model.sampling import Sampling as _Sampling
    HAVE_PYMOO = True
except Exception:       # pragma: no cover
    HAVE_PYMOO = False

    class _Sampling:
        def __init__(self):
            pass

        def do(self, problem, n_samples, **kwargs):
            return self._do(problem, n_samples, **kwargs)


class TruncatedNormalRandomSampling(_Sampling):      # operators.py:9-15
    def __init__(self, var_type=float):
        super().__init__()
        self.var_type = var_type

    def _do(self, problem, n_samples, **kwargs):
        from scipy.stats import truncnorm
        return truncnorm.rvs(-2, 2, size=(n_samples, problem.n_var)).astype(np.float32)


class NormalRandomSampling(_Sampling):               # operators.py:17-25
    def __init__(self, mu=0, std=1, var_type=float):
        super().__init__()
        self.mu = mu
        self.std = std
        self.var_type = var_type

    def _do(self, problem, n_samples, **kwargs):
        return np.random.normal(self.mu, self.std, size=(n_samples, problem.n_var))


class BinaryRandomSampling(_Sampling):               # operators.py:27-34
    def __init__(self, prob=0.5):
        super().__init__()
        self.prob = prob

    def _do(self, problem, n_samples, **kwargs):
        val = np.random.random((n_samples, problem.n_var))
        return (val < self.prob).astype(bool)


def _factories():
    """pymoo's factory functions when it imports (the reference's own path), else the restatements in ``ga.py``
    (parity unpinned: pymoo 0.4.2.1 is not on this box)."""
    if HAVE_PYMOO:
        from pymoo.factory import get_crossover, get_mutation, get_sampling
        from pymoo.operators.mixed_variable_operator import (MixedVariableCrossover, MixedVariableMutation,
                                                             MixedVariableSampling)
    else:
        from .ga import (MixedVariableCrossover, MixedVariableMutation, MixedVariableSampling, get_crossover,
                         get_mutation, get_sampling)
    return get_crossover, get_mutation, get_sampling, MixedVariableSampling, MixedVariableCrossover, MixedVariableMutation


def get_operators(config):                           # operators.py:37-82
    get_crossover, get_mutation, get_sampling, MVS, MVC, MVM = _factories()
    name = config.config
    if name in ("DeepMindBigGAN256", "DeepMindBigGAN512"):          # operators.py:38-64
        mask = ["real"] * config.dim_z + ["bool"] * config.num_classes
        sampling = MVS(mask, {"real": TruncatedNormalRandomSampling(), "bool": BinaryRandomSampling(prob=5 / 1000)})
        crossover = MVC(mask, {"real": get_crossover("real_sbx", prob=1.0, eta=3.0),
                               "bool": get_crossover("bin_hux", prob=0.2)})
        mutation = MVM(mask, {"real": get_mutation("real_pm", prob=0.5, eta=3.0),
                              "bool": get_mutation("bin_bitflip", prob=10 / 1000)})
        return dict(sampling=sampling, crossover=crossover, mutation=mutation)
    if name.split("_")[0] == "StyleGAN2":                            # operators.py:66-71
        return dict(
            sampling=NormalRandomSampling(),
            crossover=get_crossover("real_sbx", prob=1.0, eta=3.0),
            mutation=get_mutation("real_pm", prob=0.5, eta=3.0),
        )
    if name == "GPT2":                                               # operators.py:73-78
        return dict(
            sampling=get_sampling("int_random"),
            crossover=get_crossover("int_sbx", prob=1.0, eta=3.0),
            mutation=get_mutation("int_pm", prob=0.5, eta=3.0),
        )
    raise Exception("Unknown config")
