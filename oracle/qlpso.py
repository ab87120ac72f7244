"""Oracle: QLPSO backbone (PSO family, ONE particle per step, ring-neighbourhood size chosen by a tabular Q agent).
TEST INFRASTRUCTURE ONLY.

numpy restatement of `src/optimizer/qlpso_optimizer.py:7-16` (cal_reward), `:45-46` (diversity), `:48-72` (velocity with the
ring neighbourhood best), `:74-89` (init_population), `:91-126` (update) under the engine's storage contract (float32 state,
float64 arithmetic in the reference's operation order, float32 costs; the diversity scalar stays float64).  Fixed reference
configuration: NP = 30, C = 1.49618, W = 0.729844 (`:22-25`).  Pinned by `tests/golden/qlpso.npz` (the real
`QLPSO_Optimizer`, teacher-forced on float32-representable state)."""
from __future__ import annotations

import dataclasses

import numpy as np

from . import rng as orng

NP_ = 30
C_ = 1.49618
W_ = 0.729844
TAG_QLPSO = 14      # Philox draw group of the two per-step scalars (mbx_common.cuh MBX_TAG_QLPSO)
K_OF_ACTION = (4, 8, 16, 30)


@dataclasses.dataclass
class Config:
    dim: int
    maxFEs: int
    n_logpoint: int = 50
    log_interval: int = 0
    NP: int = NP_

    def __post_init__(self):
        if not self.log_interval:
            self.log_interval = self.maxFEs // self.n_logpoint


@dataclasses.dataclass
class State:
    x: np.ndarray            # [NP,D] f32
    v: np.ndarray            # [NP,D] f32
    pbest_x: np.ndarray      # [NP,D] f32
    cost: np.ndarray         # [NP] f32   current costs
    gbest_cost: np.float32
    diversity: float         # float64
    q_state: np.ndarray      # [NP] int32  last action of every particle (:120), random in 0..3 at start (:88)
    pointer: int
    fes: int
    log_index: int
    curve: list
    done: bool = False


class NumpyTape:
    def __init__(self, NP, D, noise_counts=(0, 0)):
        self.NP, self.D, self.nc = NP, D, noise_counts

    def init_pos_u(self): return np.random.rand(self.NP, self.D)
    def init_state(self): return np.random.randint(low=0, high=4, size=self.NP)
    def rand2(self): return np.random.rand(), np.random.rand()

    def noise(self, n):
        nu, nn = self.nc
        d = tuple(np.random.rand(n) for _ in range(nu)) + tuple(np.random.randn(n) for _ in range(nn))
        return d if d else None


class PhiloxTape:
    """Counter-based draws identical to mbx_qlpso.cu (generation = fes before the step)."""

    def __init__(self, NP, D, seed, lane, gen, noise_kind=0):
        self.NP, self.D = NP, D
        self.p = orng.PhiloxDraws(seed, lane, gen)
        self.noise_kind = noise_kind

    def init_pos_u(self): return orng.u01(self.p.words(orng.TAG_INIT, self.NP * self.D).reshape(self.NP, self.D, 4)[..., 0])
    def init_state(self): return (self.p.words(TAG_QLPSO, self.NP)[:, 0] & 3).astype(np.int64)

    def rand2(self):
        w = self.p.words(TAG_QLPSO, 1)[0]
        return float(orng.u01(w[0])), float(orng.u01(w[1]))

    def noise(self, n):
        from .rlepso import philox_noise
        return philox_noise(self.p, self.noise_kind, n, orng.TAG_NOISE)


def _costs(problem_eval, optimum, x32, draws):
    f = problem_eval(x32.astype(np.float64), draws)
    if optimum is not None:
        f = f - optimum
    return np.asarray(f, dtype=np.float64).astype(np.float32)


def cal_diversity(x32):
    """:45-46 on the float32 positions; numpy's own reduction order IS the definition (the kernel reproduces it)."""
    p = x32.astype(np.float64)
    return float(np.mean(np.sqrt(np.sum(np.square(p - np.mean(p, 0)), 1))))


def cal_reward(f_old, f_new, d_old, d_new):
    """:7-16.  The reference returns None when f_new is NaN; the engine treats `not (f_new < f_old)` as the >= branch."""
    if f_new < f_old:
        return 2 if d_new > d_old else 1
    return 0 if d_new > d_old else -2


def nbest_index(cost, i, action):
    """:53-70 ring neighbourhood of 4 / 8 / 16 / 30 particles centred on i, first strict minimum of the CURRENT costs."""
    n = len(cost)
    k = K_OF_ACTION[int(action)] if 0 <= int(action) < 4 else 0
    best, best_cost = None, np.inf
    for j in range(-k // 2, k // 2 + 1):
        m = (i + j) % n
        if cost[m] < best_cost:
            best_cost, best = cost[m], m
    return best


def init_population(cfg, problem_eval, optimum, lb, ub, tape):
    NP = cfg.NP
    pos = (tape.init_pos_u() * (ub - lb) + lb).astype(np.float32)
    div = cal_diversity(pos)
    c = _costs(problem_eval, optimum, pos, tape.noise(NP))
    g = np.min(c)
    qs = np.asarray(tape.init_state(), dtype=np.int32)
    S = State(x=pos.copy(), v=np.zeros_like(pos), pbest_x=pos.copy(), cost=c.copy(), gbest_cost=g, diversity=div, q_state=qs,
              pointer=0, fes=NP, log_index=1, curve=[g])
    return S, int(qs[0])


def update(cfg, S, action, problem_eval, optimum, lb, ub, tape):
    """qlpso_optimizer.py:91-126 -> (new State, next q-state, reward, done)."""
    NP = cfg.NP
    i = S.pointer
    action = int(np.squeeze(action))
    x = S.x[i].astype(np.float64); v = S.v[i].astype(np.float64); pb = S.pbest_x[i].astype(np.float64)
    nb = nbest_index(S.cost, i, action)
    r1, r2 = tape.rand2()
    if nb is None:                                                 # all neighbour costs NaN/inf: the reference would raise
        nb = i
    nbest = S.x[nb].astype(np.float64)
    nv = W_ * v + C_ * r1 * (nbest - x) + C_ * r2 * (pb - x)
    nx = np.clip(x + nv, lb, ub)
    nx32, nv32 = nx.astype(np.float32), nv.astype(np.float32)
    N = dataclasses.replace(S, x=S.x.copy(), v=S.v.copy(), pbest_x=S.pbest_x.copy(), cost=S.cost.copy(), q_state=S.q_state.copy(),
                            curve=list(S.curve))
    N.x[i], N.v[i] = nx32, nv32
    f_old = S.cost[i]
    f_new = _costs(problem_eval, optimum, nx32[None, :], tape.noise(1))[0]
    N.fes = S.fes + 1
    d_new = cal_diversity(N.x)
    reward = cal_reward(f_old, f_new, S.diversity, d_new)
    N.cost[i] = f_new
    N.diversity = d_new
    N.gbest_cost = np.float32(min(S.gbest_cost, np.min(N.cost)))
    if f_new < f_old:
        N.pbest_x[i] = nx32
    N.q_state[i] = action
    N.pointer = (i + 1) % NP
    if N.fes >= N.log_index * cfg.log_interval:
        N.log_index += 1
        N.curve.append(N.gbest_cost)
    if optimum is None:
        done = N.fes >= cfg.maxFEs
    else:
        done = bool(N.fes >= cfg.maxFEs or N.gbest_cost <= 1e-8)
    if done:
        if len(N.curve) >= cfg.n_logpoint + 1:
            N.curve[-1] = N.gbest_cost
        else:
            N.curve.append(N.gbest_cost)
    N.done = done
    return N, int(N.q_state[N.pointer]), int(reward), done


def run_episode(cfg, problem_eval, optimum, lb, ub, policy, seed, lane, noise_kind=0, max_steps=None):
    """Philox-driven episode; `policy(q_state:int, step:int) -> action int`.  -> (State, total return)."""
    S, qs = init_population(cfg, problem_eval, optimum, lb, ub, PhiloxTape(cfg.NP, cfg.dim, seed, lane, orng.GEN_INIT, noise_kind))
    R, t = 0, 0
    while not S.done and (max_steps is None or t < max_steps):
        a = policy(qs, t)
        S, qs, r, _ = update(cfg, S, a, problem_eval, optimum, lb, ub, PhiloxTape(cfg.NP, cfg.dim, seed, lane, S.fes, noise_kind))
        R += r; t += 1
    return S, R
