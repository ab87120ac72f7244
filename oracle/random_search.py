"""Oracle: Random_search baseline.  TEST INFRASTRUCTURE ONLY.

numpy restatement of `src/optimizer/random_search.py:17-59` under the engine's storage contract (positions rounded to
float32 before evaluation, costs rounded to float32).  Pinned by `tests/golden/rs.npz` (the real `Random_search.run_episode`
on an `F32Problem` proxy with `np.random.seed(5)`)."""
from __future__ import annotations

import numpy as np

from . import rng as orng


class NumpyTape:
    def __init__(self, NP, D, noise_counts=(0, 0)):
        self.NP, self.D, self.nc = NP, D, noise_counts

    def pos_u(self, it):
        return np.random.random_sample((self.NP, self.D))          # np.random.uniform(lb, ub) = lb + (ub - lb) * u

    def noise(self, it):
        nu, nn = self.nc
        d = tuple(np.random.rand(self.NP) for _ in range(nu)) + tuple(np.random.randn(self.NP) for _ in range(nn))
        return d if d else None


class PhiloxTape:
    """Counter-based draws identical to random_search_kernel (index = element, generation = iteration)."""

    def __init__(self, NP, D, seed, lane, noise_kind=0):
        self.NP, self.D, self.seed, self.lane, self.noise_kind = NP, D, seed, lane, noise_kind

    def pos_u(self, it):
        return orng.u01(orng.PhiloxDraws(self.seed, self.lane, it).words(orng.TAG_INIT, self.NP * self.D)[:, 0]).reshape(self.NP, self.D)

    def noise(self, it):
        from .rlepso import philox_noise
        return philox_noise(orng.PhiloxDraws(self.seed, self.lane, it), self.noise_kind, self.NP, orng.TAG_NOISE)


def run_episode(dim, maxFEs, n_logpoint, log_interval, problem_eval, optimum, lb, ub, tape, NP=100):
    """-> {'cost': [float32 gbest samples], 'fes': int}   (random_search.py:37-59)."""
    def population(it):
        pos = (lb + (ub - lb) * tape.pos_u(it)).astype(np.float32)
        f = problem_eval(pos.astype(np.float64), tape.noise(it))
        if optimum is not None:
            f = f - optimum
        return np.asarray(f, dtype=np.float64).astype(np.float32)

    fes = NP
    gbest = np.min(population(0))                                   # __reset (:17-22)
    cost, log_index, it = [gbest], 1, 0
    while True:
        it += 1
        c = population(it)
        fes += NP
        if gbest > np.min(c):
            gbest = np.min(c)
        if fes >= log_index * log_interval:
            log_index += 1
            cost.append(gbest)
        done = fes >= maxFEs if optimum is None else bool(gbest <= 1e-8 or fes >= maxFEs)
        if done:
            if len(cost) >= n_logpoint + 1:
                cost[-1] = gbest
            else:
                cost.append(gbest)
            break
    return {'cost': cost, 'fes': fes}
