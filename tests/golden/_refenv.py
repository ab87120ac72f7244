"""Import the UNMODIFIED reference (GMC-DRL/MetaBox) from /root/reference for golden generation.

Only `tests/golden/make_golden*.py` use this, and only in the build container: `/root/reference`
does not exist on the GPU box, so nothing under `tests/test_*.py`, `bench.py` or `smoke()` imports
this module.  Third-party packages the reference imports at package-import time but that are not
on the hot path (deap, skopt, matplotlib, openpyxl, cmaes, gym) are stubbed.
"""
import os
import sys
from unittest.mock import MagicMock

REF = os.environ.get('METABOX_REFERENCE', '/root/reference')


def activate():
    sys.dont_write_bytecode = True
    src = os.path.join(REF, 'src')
    if not os.path.isdir(src):
        raise RuntimeError(f'reference not found at {REF}')
    if src not in sys.path:
        sys.path.insert(0, src)
    for m in ['deap', 'deap.base', 'deap.creator', 'deap.tools', 'deap.cma', 'deap.algorithms', 'skopt',
              'matplotlib', 'matplotlib.pyplot', 'openpyxl', 'cmaes', 'gym', 'gym.spaces']:
        sys.modules.setdefault(m, MagicMock())
    return src


class F32Problem:
    """Proxy handed to the reference optimizers so that they run under the engine's storage contract
    without touching reference code: positions are rounded to float32 before evaluation and the
    returned cost `eval(x) - optimum` is a float32-representable float64.  `optimum` is reported as 0
    (or None when the wrapped problem has none) so the reference's `eval(x) - problem.optimum` and
    its `gbest <= 1e-8` termination test keep their meaning."""

    def __init__(self, inner, inplace=False):
        """inplace=True additionally writes the rounded positions back into the caller's array, so that what the reference optimizer
        STORES afterwards (trial vectors, walk samples) is the float32-representable point that was evaluated."""
        import numpy as np
        self._np = np
        self.inner = inner
        self.inplace = inplace
        self.dim, self.lb, self.ub = inner.dim, inner.lb, inner.ub
        self.optimum = None if inner.optimum is None else 0.0
        self.T1 = 0

    def reset(self):
        self.inner.reset()
        self.T1 = 0

    def __str__(self):
        return str(self.inner)

    def eval(self, x):
        np = self._np
        x32 = np.asarray(x, dtype=np.float64).astype(np.float32).astype(np.float64)
        if self.inplace and isinstance(x, np.ndarray) and x.flags.writeable and x.dtype == np.float64:
            x[...] = x32
        f = self.inner.eval(x32)
        if self.inner.optimum is not None:
            f = f - self.inner.optimum
        return np.asarray(f, dtype=np.float64).astype(np.float32).astype(np.float64) if np.ndim(f) else \
            float(np.float32(f))
