"""Oracle: DEDQN backbone (DE family, ONE individual per step, landscape-feature state).  TEST INFRASTRUCTURE ONLY.

numpy restatement of `src/optimizer/dedqn_optimizer.py:79-206` (+ `operators/mutate.py:5-9,36-39,88-91,178-181`,
`operators/crossover.py:6-18`, `operators/boundary_control.py:5-9`; the four features are `oracle/landscape.py`) under the engine's
storage contract: population / costs are float32, arithmetic is float64 on the up-cast state in the reference's operation order, the
trial is rounded to float32 before evaluation and every cost `eval(x) - optimum` is rounded to float32.

Reference quirks reproduced on purpose:
  * `__cal_feature` (:123-136) draws a random walk of `rwsteps = NP` points inside the population's bounding box but evaluates the
    POPULATION, not the walk (`problem.eval(self.__population)`): the features pair walk positions with population costs; it charges
    `rwsteps` evaluations, and `update` charges `NP` for its single trial (:172) -- 200 charged evaluations per step;
  * `__gbest` starts as a *view* of row argmin(cost) of the population (:145) and follows in-place writes to that row (:174) until it
    is re-bound to a trial array (:178).  Observable only through `best_2_single` when a trial EQUAL in cost to the best row replaces it;
  * `generate_random_int_single` redraws all indices until none equals the pointer; the indices may repeat among themselves;
  * termination looks at `cost.min()`, the logged value is `gbest_cost` (identical by construction).

Pinned by `tests/golden/dedqn.npz` (the real `DEDQN_Optimizer` run on F32Problem proxies, tests/golden/make_golden_dedqn.py).
"""
from __future__ import annotations

import dataclasses

import numpy as np

from . import landscape as ola

F_, CR_ = 0.5, 0.5          # dedqn_optimizer.py:107-108
N_COLS = (3, 3, 4)          # indices drawn by rand_1_single / cur_to_rand_1_single / best_2_single


@dataclasses.dataclass
class Config:
    dim: int
    maxFEs: int
    n_logpoint: int = 50
    log_interval: int = 0
    NP: int = 100            # :106 (rwsteps = NP, :109)

    def __post_init__(self):
        if not self.log_interval:
            self.log_interval = self.maxFEs // self.n_logpoint


@dataclasses.dataclass
class State:
    pop: np.ndarray          # [NP,D] f32
    cost: np.ndarray         # [NP]   f32
    survival: np.ndarray     # [NP]   f64 (counts)
    gbest_row: int           # >= 0: gbest is still a view of this population row ; -1: own values below
    gbest: np.ndarray        # [D] f32 (ignored while gbest_row >= 0)
    gbest_cost: np.float32
    pointer: int
    fes: int
    log_index: int
    curve: list
    done: bool = False

    def best(self):
        return self.pop[self.gbest_row] if self.gbest_row >= 0 else self.gbest

    def copy(self):
        kw = {f.name: (getattr(self, f.name).copy() if isinstance(getattr(self, f.name), np.ndarray) else getattr(self, f.name))
              for f in dataclasses.fields(self)}
        kw['curve'] = list(self.curve)
        return State(**kw)


class NumpyTape:
    """Reference draw order on the global np.random; `rec` keeps what was drawn so that the engine can replay it."""

    def __init__(self, NP, D, noise_counts=(0, 0)):
        self.NP, self.D, self.nc = NP, D, noise_counts
        self.rec = {}

    def init_u(self):
        self.rec['init_u'] = np.random.rand(self.NP, self.D)                                # :139
        return self.rec['init_u']

    def walk_u(self):
        self.rec['walk_u'] = np.stack([np.random.rand(self.D) for _ in range(self.NP)])     # :79-90: start point, then NP - 1 moves
        return self.rec['walk_u']

    def r(self, cols, pointer):
        r = np.random.randint(low=0, high=self.NP, size=cols)                               # mutate.py:5-9
        while pointer in r:
            r = np.random.randint(low=0, high=self.NP, size=cols)
        self.rec['r'] = np.concatenate([r, np.zeros(4 - cols, dtype=r.dtype)])
        return r

    def crossover(self):
        self.rec['jrand'] = int(np.random.randint(self.D, size=1)[0])                       # crossover.py:11
        self.rec['cr_u'] = np.random.rand(1, self.D)[0]                                     # crossover.py:14
        return self.rec['jrand'], self.rec['cr_u']

    def noise(self, n, key):
        nu, nn = self.nc
        d = tuple(np.random.rand(n) for _ in range(nu)) + tuple(np.random.randn(n) for _ in range(nn))
        self.rec[key] = d
        return d if d else None


class ArrayTape:
    """Replays recorded draws (a NumpyTape's `rec`, or the tape the batched engine drew on the device)."""

    def __init__(self, rec):
        self.rec = rec

    def init_u(self): return np.asarray(self.rec['init_u'], dtype=np.float64)
    def walk_u(self): return np.asarray(self.rec['walk_u'], dtype=np.float64)
    def r(self, cols, pointer): return np.asarray(self.rec['r'])[:cols]
    def crossover(self): return int(self.rec['jrand']), np.asarray(self.rec['cr_u'], dtype=np.float64)
    def noise(self, n, key): return self.rec.get(key) or None


def f32(a):
    return np.asarray(a, dtype=np.float64).astype(np.float32)


def random_walk(pop, walk_u):
    """dedqn_optimizer.py:79-90 with the NP uniform vectors given: cumulative sum modulo 1, scaled into the population's bounding box."""
    pmin, pmax = np.min(pop, axis=0), np.max(pop, axis=0)
    pts, cur = [], None
    for k in range(walk_u.shape[0]):
        cur = walk_u[k] if k == 0 else (cur + walk_u[k]) % 1
        pts.append(cur.tolist())
    return pmin + (pmax - pmin) * np.array(pts)


def cal_feature(cfg, pop32, ev, optimum, tape, key):
    """:123-136 -> (state [4] f64, population costs as evaluated for the features)."""
    pop = pop32.astype(np.float64)
    samples = random_walk(pop, tape.walk_u())
    f = ev(pop, tape.noise(cfg.NP, key))
    c = f32(f if optimum is None else f - optimum).astype(np.float64)
    return np.array([ola.cal_fdc(samples, c), ola.cal_rie(c), ola.cal_acf(c), ola.cal_nop(samples, c)]), c


def cal_reward(survival, pointer):
    """:93-101"""
    reward = 0
    for i in range(len(survival)):
        if i == pointer:
            if survival[i] == 1:
                reward += 1
        else:
            reward += 1 / survival[i]
    return reward / len(survival)


def init_population(cfg: Config, ev, optimum, lb, ub, tape):
    """:138-152.  `ev(x64 [n,D], noise_draws) -> f64 [n]` evaluates float32-representable positions."""
    pop = f32(tape.init_u() * (ub - lb) + lb)
    f = ev(pop.astype(np.float64), tape.noise(cfg.NP, 'noise_init'))
    cost = f32(f if optimum is None else f - optimum)
    S = State(pop=pop, cost=cost, survival=np.ones(cfg.NP), gbest_row=int(np.argmin(cost)), gbest=pop[int(np.argmin(cost))].copy(),
              gbest_cost=np.float32(cost.min()), pointer=0, fes=cfg.NP, log_index=1, curve=[np.float32(cost.min())])
    state, _ = cal_feature(cfg, S.pop, ev, optimum, tape, 'noise_feat')
    S.fes += cfg.NP
    return S, state


def update(cfg: Config, S: State, action, ev, optimum, lb, ub, tape):
    """:154-206 -> (new State, state [4], reward, done)."""
    S = S.copy()
    if action not in (0, 1, 2):
        raise ValueError(f'action error: {action}')
    x = S.pop.astype(np.float64)
    p = S.pointer
    r = tape.r(N_COLS[action], p)
    if action == 0:
        v = x[r[0]] + F_ * (x[r[1]] - x[r[2]])                                               # mutate.py:36-39
    elif action == 1:
        v = x[p] + F_ * (x[r[0]] - x[p] + x[r[1]] - x[r[2]])                                 # mutate.py:178-181
    else:
        v = S.best().astype(np.float64) + F_ * (x[r[0]] - x[r[1]] + x[r[2]] - x[r[3]])       # mutate.py:88-91
    v = np.clip(v, lb, ub)                                                                   # boundary_control.py:5-9
    jrand, cr_u = tape.crossover()
    u = np.where(cr_u < CR_, v, x[p])                                                        # crossover.py:6-18
    u[jrand] = v[jrand]
    u32 = f32(u)
    f = ev(u32.astype(np.float64)[None, :], tape.noise(1, 'noise_trial'))
    u_cost = f32(f if optimum is None else f - optimum)[0]
    S.fes += cfg.NP
    if u_cost <= S.cost[p]:
        S.pop[p] = u32
        S.cost[p] = u_cost
        S.survival[p] = 1
        if u_cost < S.gbest_cost:
            S.gbest, S.gbest_row, S.gbest_cost = u32.copy(), -1, np.float32(u_cost)
    else:
        S.survival[p] += 1
    state, _ = cal_feature(cfg, S.pop, ev, optimum, tape, 'noise_feat')
    S.fes += cfg.NP
    if S.fes >= S.log_index * cfg.log_interval:
        S.log_index += 1
        S.curve.append(S.gbest_cost)
    reward = cal_reward(S.survival, p)
    S.done = bool(S.fes >= cfg.maxFEs) if optimum is None else bool(S.fes >= cfg.maxFEs or S.cost.min() <= 1e-8)
    if S.done:
        if len(S.curve) >= cfg.n_logpoint + 1:
            S.curve[-1] = S.gbest_cost
        else:
            S.curve.append(S.gbest_cost)
    S.pointer = (p + 1) % cfg.NP
    return S, state, reward, S.done
