"""Oracle: RL_PSO backbone (PSO family, ONE particle per step).  TEST INFRASTRUCTURE ONLY.

numpy restatement of `src/optimizer/rl_pso_optimizer.py:32-62` (init_population), `:64-65` (state), `:79-148` (update)
under the engine's storage contract (float32 state, float64 arithmetic in the reference's operation order, float32
costs).  Fixed reference configuration: NP = 100, c = 2.05, w = 0.9 decaying by 0.5 / (maxFEs / NP) per step (`:10-24`).
Pinned by `tests/golden/rl_pso.npz` (the real `RL_PSO_Optimizer`, teacher-forced on float32-representable state)."""
from __future__ import annotations

import dataclasses

import numpy as np

from . import rng as orng

NP_ = 100
C_ = 2.05
TAG_RLPSO = 13      # Philox draw group of the per-step rand1 (mbx_common.cuh MBX_TAG_RLPSO)


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
    c_cost: np.ndarray       # [NP] f32
    pbest: np.ndarray        # [NP] f32
    gbest_x: np.ndarray      # [D] f32
    gbest_val: np.float32
    gbest_idx: int
    w: float
    max_cost: np.float32     # max of the initial costs (:44)
    cur_index: int
    fes: int
    log_index: int
    curve: list
    done: bool = False


class NumpyTape:
    def __init__(self, NP, D, noise_counts=(0, 0)):
        self.NP, self.D, self.nc = NP, D, noise_counts

    def init_pos_u(self): return np.random.random_sample((self.NP, self.D))
    def init_vel_u(self): return np.random.random_sample((self.NP, self.D))
    def rand1(self): return np.random.rand()

    def noise(self, n):
        nu, nn = self.nc
        d = tuple(np.random.rand(n) for _ in range(nu)) + tuple(np.random.randn(n) for _ in range(nn))
        return d if d else None


class PhiloxTape:
    """Counter-based draws identical to mbx_rlpso.cu (generation = fes before the step, like DE-DDQN)."""

    def __init__(self, NP, D, seed, lane, gen, noise_kind=0):
        self.NP, self.D = NP, D
        self.p = orng.PhiloxDraws(seed, lane, gen)
        self.noise_kind = noise_kind

    def init_pos_u(self): return orng.u01(self.p.words(orng.TAG_INIT, self.NP * self.D).reshape(self.NP, self.D, 4)[..., 0])
    def init_vel_u(self): return orng.u01(self.p.words(orng.TAG_INIT, self.NP * self.D).reshape(self.NP, self.D, 4)[..., 1])
    def rand1(self): return float(orng.u01(self.p.words(TAG_RLPSO, 1)[0, 0]))

    def noise(self, n):
        from .rlepso import philox_noise
        return philox_noise(self.p, self.noise_kind, n, orng.TAG_NOISE)


def _costs(problem_eval, optimum, x32, draws):
    f = problem_eval(x32.astype(np.float64), draws)
    if optimum is not None:
        f = f - optimum
    return np.asarray(f, dtype=np.float64).astype(np.float32)


def get_state(S):
    """:64-65 -> float64 [2 D] = gbest_position | current_position[cur_index]."""
    return np.concatenate([S.gbest_x.astype(np.float64), S.x[S.cur_index].astype(np.float64)])


def init_population(cfg, problem_eval, optimum, lb, ub, tape):
    NP, D = cfg.NP, cfg.dim
    vmax = 0.1 * (ub - lb)
    pos = (lb + (ub - lb) * tape.init_pos_u()).astype(np.float32)
    vel = (-vmax + (vmax - (-vmax)) * tape.init_vel_u()).astype(np.float32)
    c = _costs(problem_eval, optimum, pos, tape.noise(NP))
    gi = int(np.argmin(c))
    S = State(x=pos.copy(), v=vel.copy(), pbest_x=pos.copy(), c_cost=c.copy(), pbest=c.copy(), gbest_x=pos[gi].copy(),
              gbest_val=c[gi], gbest_idx=gi, w=0.9, max_cost=np.max(c), cur_index=0, fes=NP, log_index=1, curve=[c[gi]])
    return S, get_state(S)


def update(cfg, S, action, problem_eval, optimum, lb, ub, tape):
    """rl_pso_optimizer.py:79-148 -> (new State, state [2D], reward, done).  `action`: float32 scalar (shape (1,))."""
    NP = cfg.NP
    vmax = 0.1 * (ub - lb)
    w = S.w - 0.5 / (cfg.maxFEs / NP)
    rand1 = tape.rand1()
    rand2 = np.float32(np.squeeze(np.asarray(action, dtype=np.float32)))
    j = S.cur_index
    v = S.v[j].astype(np.float64); x = S.x[j].astype(np.float64)
    pb = S.pbest_x[j].astype(np.float64); gx = S.gbest_x.astype(np.float64)
    pre_cost = S.c_cost[j]
    c2 = np.float32(C_) * rand2                                    # numpy-2: python float x float32 scalar -> float32
    nv = w * v + C_ * rand1 * (pb - x) + float(c2) * (gx - x)
    nv = np.clip(nv, -vmax, vmax)
    nx = np.clip(x + nv, lb, ub)
    nx32, nv32 = nx.astype(np.float32), nv.astype(np.float32)
    N = dataclasses.replace(S, x=S.x.copy(), v=S.v.copy(), pbest_x=S.pbest_x.copy(), c_cost=S.c_cost.copy(), pbest=S.pbest.copy(),
                            gbest_x=S.gbest_x.copy(), curve=list(S.curve), w=w)
    N.x[j], N.v[j] = nx32, nv32
    new_cost = _costs(problem_eval, optimum, nx32[None, :], tape.noise(1))[0]
    N.fes = S.fes + 1
    N.c_cost[j] = new_cost
    if new_cost < N.pbest[j]:
        N.pbest[j] = new_cost; N.pbest_x[j] = nx32
    if new_cost < N.gbest_val:
        N.gbest_val = new_cost; N.gbest_x = nx32.copy(); N.gbest_idx = j
    if optimum is None:
        done = N.fes >= cfg.maxFEs
    else:
        done = bool(N.fes >= cfg.maxFEs or N.gbest_val <= 1e-8)
    if N.fes >= N.log_index * cfg.log_interval:
        N.log_index += 1
        N.curve.append(N.gbest_val)
    with np.errstate(divide='ignore', invalid='ignore'):
        reward = (np.float64(pre_cost) - np.float64(new_cost)) / (np.float64(S.max_cost) - np.float64(N.gbest_val))
    N.cur_index = (j + 1) % NP
    if done:
        if len(N.curve) >= cfg.n_logpoint + 1:
            N.curve[-1] = N.gbest_val
        else:
            N.curve.append(N.gbest_val)
    N.done = done
    return N, get_state(N), float(reward), done
