"""Mirror of the reference's operators.py.

The three ``Sampling`` subclasses (operators.py:9-34) are pure numpy/scipy and
are restated here (with ``float``/``bool`` for the ``np.float``/``np.bool``
aliases that numpy>=1.24 removed).  SBX / PM / HUX / bit-flip live in pymoo
0.4.2.1, which is not installable offline: ``get_operators`` wires them exactly
as operators.py:37-82 when pymoo is importable and raises otherwise — their
arithmetic is NOT re-implemented here (parity would be unpinned).
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


def get_operators(config):                           # operators.py:37-82
    name = config.config
    if name.split("_")[0] == "StyleGAN2":
        if not HAVE_PYMOO:
            raise ImportError("pymoo==0.4.2.1 is required for real_sbx / real_pm (operators.py:69-70)")
        from pymoo.factory import get_crossover, get_mutation
        return dict(
            sampling=NormalRandomSampling(),
            crossover=get_crossover("real_sbx", prob=1.0, eta=3.0),
            mutation=get_mutation("real_pm", prob=0.5, eta=3.0),
        )
    if name in ("DeepMindBigGAN256", "DeepMindBigGAN512", "GPT2"):
        raise NotImplementedError(f"{name}: 'next' row of SURVEY.md §8(f)")
    raise Exception("Unknown config")
