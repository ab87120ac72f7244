"""Oracle: GLEET backbone (PSO family, whole swarm per step, 27 features per particle).  TEST INFRASTRUCTURE ONLY.

numpy restatement of `src/optimizer/gleet_optimizer.py:34-60` (initialize_particles), `:76-112` (init_population),
`:128-158` (observe / gp_cat), `:193-314` (update) under the engine's storage contract: population state (positions,
velocities, pbest positions, costs, carried features) lives in float32, every arithmetic chain runs in float64 on the
up-cast state in the reference's operation order, costs are rounded to float32 at the evaluation boundary and all
comparisons are made on those float32 costs.  Fixed reference configuration: ps = 100, c = 4.1, w 0.9 decaying by
0.5 / (maxFEs / ps) per step, clipping border, 'direct' reward x 100 (`:10-32`).

Pinned by `tests/golden/gleet.npz`: the real `GLEET_Optimizer.init_population / update`, teacher-forced on
float32-representable state with `np.random.seed(s)`, must give the same next state.
"""
from __future__ import annotations

import dataclasses

import numpy as np

from . import rng as orng

NP_ = 100           # gleet_optimizer.py:21
C_ = 4.1            # :16
REWARD_SCALE = 100  # :19
N_FEAT = 9
TAG_GLEET_P = 12    # Philox draw group of the per-particle (rand1, rand2) pair (mbx_common.cuh MBX_TAG_GLEET_P)


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
    x: np.ndarray              # [NP,D] f32  current_position
    v: np.ndarray              # [NP,D] f32  velocity
    pbest_x: np.ndarray        # [NP,D] f32
    c_cost: np.ndarray         # [NP]   f32
    pbest: np.ndarray          # [NP]   f32
    pni: np.ndarray            # [NP]   i32  per_no_improve
    gbest_x: np.ndarray        # [D]    f32
    gbest_val: np.float32
    gbest_idx: int
    no_improve: int
    w: float                   # inertia (float64, decays)
    max_cost: np.float32       # min of the initial costs (`self.max_cost = np.min(c_cost)`, :49)
    pbest_feature: np.ndarray  # [NP,9] f32
    gbest_feature: np.ndarray  # [9]    f32
    fes: int
    log_index: int
    curve: list
    done: bool = False


class NumpyTape:
    """Draws from the global np.random in the reference's call order."""

    def __init__(self, NP, D, noise_counts=(0, 0)):
        self.NP, self.D, self.nc = NP, D, noise_counts

    def init_pos_u(self): return np.random.random_sample((self.NP, self.D))      # np.random.uniform = low + (high-low) * u
    def init_vel_u(self): return np.random.random_sample((self.NP, self.D))
    def rand1(self): return np.random.rand(self.NP, 1)[:, 0]
    def rand2(self): return np.random.rand(self.NP, 1)[:, 0]

    def noise(self, which):
        nu, nn = self.nc
        d = tuple(np.random.rand(self.NP) for _ in range(nu)) + tuple(np.random.randn(self.NP) for _ in range(nn))
        return d if d else None


class ArrayTape:
    def __init__(self, rec):
        self.rec = rec

    def __getattr__(self, k):
        if k == 'noise':
            return lambda which: self.rec.get(f'noise{which}') or None
        if k == 'rec':
            raise AttributeError(k)
        return lambda: self.rec[k]


class PhiloxTape:
    """Counter-based draws identical to the CUDA production path (mbx_gleet.cu)."""

    def __init__(self, NP, D, seed, lane, gen, noise_kind=0):
        self.NP, self.D = NP, D
        self.p = orng.PhiloxDraws(seed, lane, gen)
        self.noise_kind = noise_kind

    def init_pos_u(self): return orng.u01(self.p.words(orng.TAG_INIT, self.NP * self.D).reshape(self.NP, self.D, 4)[..., 0])
    def init_vel_u(self): return orng.u01(self.p.words(orng.TAG_INIT, self.NP * self.D).reshape(self.NP, self.D, 4)[..., 1])
    def rand1(self): return orng.u01(self.p.words(TAG_GLEET_P, self.NP)[:, 0])
    def rand2(self): return orng.u01(self.p.words(TAG_GLEET_P, self.NP)[:, 1])

    def noise(self, which):
        from .rlepso import philox_noise
        return philox_noise(self.p, self.noise_kind, self.NP, orng.TAG_NOISE + which)


def _costs(problem_eval, optimum, x32, noise_draws):
    f = problem_eval(x32.astype(np.float64), noise_draws)
    if optimum is not None:
        f = f - optimum
    return np.asarray(f, dtype=np.float64).astype(np.float32)


def observe(cfg, x, pbest_x, gbest_x, c_cost, pbest, gbest_val, pni, no_improve, fes, max_cost, lb, ub):
    """gleet_optimizer.py:128-154 on the up-cast float32 state -> float64 [NP, 9]."""
    NP = cfg.NP
    x = x.astype(np.float64); px = pbest_x.astype(np.float64); gx = gbest_x.astype(np.float64)
    cc = c_cost.astype(np.float64); pb = pbest.astype(np.float64)
    gb, mc = float(gbest_val), float(max_cost)
    max_step = cfg.maxFEs // NP
    max_dist = np.sqrt((ub - lb) ** 2 * cfg.dim)
    fea0 = cc / mc
    fea1 = (cc - gb) / mc
    fea2 = (cc - pb) / mc
    fea3 = np.full(NP, (cfg.maxFEs - fes) / cfg.maxFEs)
    fea4 = pni.astype(np.float64) / max_step
    fea5 = np.full(NP, no_improve / max_step)
    fea6 = np.sqrt(np.sum((x - gx[None, :]) ** 2, axis=-1)) / max_dist
    fea7 = np.sqrt(np.sum((x - px) ** 2, axis=-1)) / max_dist
    pv = px - x
    gv = gx[None, :] - x
    with np.errstate(invalid='ignore', divide='ignore'):
        fea8 = np.sum(pv * gv, axis=-1) / ((np.sqrt(np.sum(pv ** 2, axis=-1)) * np.sqrt(np.sum(gv ** 2, axis=-1))) + 1e-5)
    fea8 = np.where(np.isnan(fea8), np.zeros_like(fea8), fea8)
    return np.stack([fea0, fea1, fea2, fea3, fea4, fea5, fea6, fea7, fea8], axis=-1)


def _full_state(feat, pbest_feature, gbest_feature):
    """:110-112 / :299-300: [NP, 9 + 18] = observe | pbest_feature | gbest_feature (repeated)."""
    NP = feat.shape[0]
    return np.concatenate([feat, pbest_feature.astype(np.float64),
                           np.repeat(gbest_feature.astype(np.float64)[None, :], NP, axis=0)], axis=-1)


def init_population(cfg, problem_eval, optimum, lb, ub, tape):
    """gleet_optimizer.py:76-112 (+ initialize_particles :34-60) -> (State, state [NP,27] float64)."""
    NP, D = cfg.NP, cfg.dim
    vmax = 0.1 * (ub - lb)
    pos = (lb + (ub - lb) * tape.init_pos_u()).astype(np.float32)
    vel = (-vmax + (vmax - (-vmax)) * tape.init_vel_u()).astype(np.float32)
    c = _costs(problem_eval, optimum, pos, tape.noise(0))
    gi = int(np.argmin(c))
    pni = np.zeros(NP, dtype=np.int32)
    feat = observe(cfg, pos, pos, pos[gi], c, c, c[gi], pni, 0, NP, c[gi], lb, ub)
    pf = feat.astype(np.float32)
    S = State(x=pos.copy(), v=vel.copy(), pbest_x=pos.copy(), c_cost=c.copy(), pbest=c.copy(), pni=pni,
              gbest_x=pos[gi].copy(), gbest_val=c[gi], gbest_idx=gi, no_improve=0, w=0.9, max_cost=c[gi],
              pbest_feature=pf, gbest_feature=pf[gi].copy(), fes=NP, log_index=1, curve=[c[gi]])
    return S, _full_state(feat, S.pbest_feature, S.gbest_feature)


def update(cfg, S, action, problem_eval, optimum, lb, ub, tape, aux=None):
    """gleet_optimizer.py:193-314 -> (new State, state [NP,27] float64, reward, done).  `action` float32 [NP] in [0,1]."""
    NP, D = cfg.NP, cfg.dim
    vmax = 0.1 * (ub - lb)
    pre_gbest = S.gbest_val
    w = S.w - 0.5 / (cfg.maxFEs / NP)                                              # :201-202
    r1 = tape.rand1()[:, None]
    r2 = tape.rand2()[:, None]
    a = np.asarray(action, dtype=np.float32)[:, None]
    x = S.x.astype(np.float64); v = S.v.astype(np.float64); px = S.pbest_x.astype(np.float64)
    gx = S.gbest_x.astype(np.float64)
    # numpy-2 promotion: python floats are weak, `self.c * action` and `self.c * (1 - action)` stay float32 (:212-213)
    ca = (np.float32(C_) * a).astype(np.float32)
    cb = (np.float32(C_) * (np.float32(1) - a)).astype(np.float32)
    nv = w * v + ca * r1 * (px - x) + cb * r2 * (gx[None, :] - x)
    nv = np.clip(nv, -vmax, vmax)
    nx = np.clip(x + nv, lb, ub)                                                   # 'clipping' border (:219-221)
    nx32 = nx.astype(np.float32); nv32 = nv.astype(np.float32)
    fes = S.fes + NP
    nc = _costs(problem_eval, optimum, nx32, tape.noise(0))
    better = nc < S.pbest
    bi = int(np.argmin(nc)); bv = nc[bi]
    gb_better = bool(bv < S.gbest_val)
    gbest_val = bv if gb_better else S.gbest_val
    no_improve = 0 if gbest_val < S.gbest_val else S.no_improve + 1               # :264-267
    pni = np.where(nc < S.c_cost, 0, S.pni + 1).astype(np.int32)                  # :270-273
    N = State(x=nx32, v=nv32, pbest_x=np.where(better[:, None], nx32, S.pbest_x), c_cost=nc,
              pbest=np.where(better, nc, S.pbest), pni=pni,
              gbest_x=nx32[bi].copy() if gb_better else S.gbest_x.copy(), gbest_val=gbest_val,
              gbest_idx=bi if gb_better else S.gbest_idx, no_improve=no_improve, w=w, max_cost=S.max_cost,
              pbest_feature=S.pbest_feature, gbest_feature=S.gbest_feature, fes=fes, log_index=S.log_index,
              curve=list(S.curve))
    if optimum is None:
        done = fes >= cfg.maxFEs
    else:
        done = bool(fes >= cfg.maxFEs or gbest_val <= 1e-8)
    reward = (float(pre_gbest) - float(gbest_val)) / float(S.max_cost) * REWARD_SCALE   # 'direct' (:160-163, :296)
    feat = observe(cfg, N.x, N.pbest_x, N.gbest_x, N.c_cost, N.pbest, N.gbest_val, N.pni, N.no_improve, fes, N.max_cost, lb, ub)
    f32 = feat.astype(np.float32)
    N.pbest_feature = np.where(pni[:, None] == 0, f32, S.pbest_feature)            # :302
    N.gbest_feature = f32[N.gbest_idx].copy() if no_improve == 0 else S.gbest_feature.copy()   # :304-305
    if fes >= N.log_index * cfg.log_interval:
        N.log_index += 1
        N.curve.append(gbest_val)
    if done:
        if len(N.curve) >= cfg.n_logpoint + 1:
            N.curve[-1] = gbest_val
        else:
            N.curve.append(gbest_val)
    N.done = done
    if aux is not None:
        aux.update(better=better, cost=nc, feat=feat)
    return N, _full_state(feat, N.pbest_feature, N.gbest_feature), reward, done
