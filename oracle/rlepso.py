"""Oracle: RLEPSO backbone (PSO family) -- init_population / update.  TEST INFRASTRUCTURE ONLY.

numpy restatement of `src/optimizer/rlepso_optimizer.py:39-263` under the engine's storage
contract: population state lives in float32 (x, v, pbest_x, costs), every arithmetic chain runs in
float64 on the up-cast state in the reference's operation order, costs are rounded to float32 at
the evaluation boundary and all comparisons (pbest `<`, gbest `<`, argmin, patience) are made on
those float32 costs.  Random numbers come from a *tape* object so that the same draws can be fed to
the CUDA kernel (tape mode) or regenerated from Philox counters (production mode).

Pinned by `tests/golden/rlepso_*.npz`: the real reference `RLEPSO_Optimizer.update`, started from
float32-representable state with `np.random.seed(s)`, must give the same next state.
"""
from __future__ import annotations

import dataclasses

import numpy as np

from . import rng as orng

NP_ = 100          # rlepso_optimizer.py:11
N_GROUP = 5        # rlepso_optimizer.py:27
ACTION_DIM = 35


@dataclasses.dataclass
class Config:
    dim: int
    maxFEs: int
    n_logpoint: int = 50
    log_interval: int = 0
    NP: int = NP_

    def __post_init__(self):
        if not self.log_interval:
            self.log_interval = self.maxFEs // self.n_logpoint      # config.py:102


def pci_table(NP):
    """rlepso_optimizer.py:24-25."""
    idx = np.array(list(range(NP)))
    return 0.05 + 0.45 * np.exp(10 * idx / (NP - 1)) / (np.exp(10) - 1)


@dataclasses.dataclass
class State:
    x: np.ndarray            # [NP,D] f32  current_position
    v: np.ndarray            # [NP,D] f32  velocity
    pbest_x: np.ndarray      # [NP,D] f32
    c_cost: np.ndarray       # [NP]   f32
    pbest: np.ndarray        # [NP]   f32
    pni: np.ndarray          # [NP]   i32  per_no_improve
    gbest_x: np.ndarray      # [D]    f32
    gbest_val: np.float32
    gbest_idx: int
    no_improve: int
    fes: int
    log_index: int
    curve: list              # gbest samples (float32 values)
    done: bool = False

    def copy(self):
        return dataclasses.replace(self, x=self.x.copy(), v=self.v.copy(), pbest_x=self.pbest_x.copy(),
                                   c_cost=self.c_cost.copy(), pbest=self.pbest.copy(), pni=self.pni.copy(),
                                   gbest_x=self.gbest_x.copy(), curve=list(self.curve))


# ----------------------------------------------------------------------------------- tapes
class NumpyTape:
    """Draws from the global np.random in the reference's call order (`NP`, `D` fixed)."""

    def __init__(self, NP, D, noise_counts=(0, 0)):
        self.NP, self.D, self.nc = NP, D, noise_counts
        self.rec = {}

    def _keep(self, k, a):
        self.rec[k] = a
        return a

    def init_pos_u(self):  return self._keep('init_pos_u', np.random.random_sample((self.NP, self.D)))
    def init_vel_u(self):  return self._keep('init_vel_u', np.random.random_sample((self.NP, self.D)))
    def rand1(self):       return self._keep('rand1', np.random.rand(self.NP, 1)[:, 0])
    def rand2(self):       return self._keep('rand2', np.random.rand(self.NP, 1)[:, 0])
    def clpso_u(self):     return self._keep('clpso_u', np.random.rand(self.NP, self.D))
    def tour(self):        return self._keep('tour', np.random.randint(low=0, high=self.NP, size=(self.NP, self.D, 2)))
    def fdr_u(self):       return self._keep('fdr_u', np.random.rand(self.NP, self.D))
    def reinit_u(self):    return self._keep('reinit_u', np.random.rand(self.NP))
    def reinit_pos_u(self): return self._keep('reinit_pos_u', np.random.random_sample((self.NP, self.D)))
    def reinit_vel_u(self): return self._keep('reinit_vel_u', np.random.random_sample((self.NP, self.D)))

    def noise(self, which):
        nu, nn = self.nc
        d = tuple(np.random.rand(self.NP) for _ in range(nu)) + tuple(np.random.randn(self.NP) for _ in range(nn))
        self.rec[f'noise{which}'] = d
        return d if d else None


class ArrayTape:
    """Replays recorded arrays (dict from NumpyTape.rec or hand-made)."""

    def __init__(self, rec):
        self.rec = rec

    def __getattr__(self, k):
        if k == 'noise':
            return lambda which: self.rec.get(f'noise{which}') or None
        if k in ('rec',):
            raise AttributeError(k)
        return lambda: self.rec[k]


class PhiloxTape:
    """Counter-based draws identical to the CUDA production path (mbx_philox.cuh)."""

    def __init__(self, NP, D, seed, lane, gen, noise_kind=0):
        self.NP, self.D = NP, D
        self.p = orng.PhiloxDraws(seed, lane, gen)
        self.noise_kind = noise_kind
        self._pd = None
        self._pp = None
        self._ri = None
        self._in = None

    def _PD(self):
        if self._pd is None:
            self._pd = self.p.words(orng.TAG_PD, self.NP * self.D).reshape(self.NP, self.D, 4)
        return self._pd

    def _P(self):
        if self._pp is None:
            self._pp = self.p.words(orng.TAG_P, self.NP)
        return self._pp

    def _RI(self):
        if self._ri is None:
            self._ri = self.p.words(orng.TAG_REINIT, self.NP * self.D).reshape(self.NP, self.D, 4)
        return self._ri

    def _IN(self):
        if self._in is None:
            self._in = self.p.words(orng.TAG_INIT, self.NP * self.D).reshape(self.NP, self.D, 4)
        return self._in

    def init_pos_u(self):  return orng.u01(self._IN()[..., 0])
    def init_vel_u(self):  return orng.u01(self._IN()[..., 1])
    def rand1(self):       return orng.u01(self._P()[:, 0])
    def rand2(self):       return orng.u01(self._P()[:, 1])
    def reinit_u(self):    return orng.u01(self._P()[:, 2])
    def clpso_u(self):     return orng.u01(self._PD()[..., 0])
    def tour(self):        return orng.bounded(self._PD()[..., 1:3], self.NP)
    def fdr_u(self):       return orng.u01(self._PD()[..., 3])
    def reinit_pos_u(self): return orng.u01(self._RI()[..., 0])
    def reinit_vel_u(self): return orng.u01(self._RI()[..., 1])

    def noise(self, which):
        return philox_noise(self.p, self.noise_kind, self.NP, orng.TAG_NOISE + which)


def philox_noise(p, kind, n, tag):
    """Noise draws for n candidates from Philox words (same mapping as mbx_bbob.cuh)."""
    if kind == 0:
        return None
    w = p.words(tag, n)
    if kind == 1:                                   # gauss : 1 normal
        n0, _ = orng.box_muller(w[:, 0], w[:, 1])
        return (n0,)
    if kind == 2:                                   # uniform: 2 uniforms
        return (orng.u01(w[:, 0]), orng.u01(w[:, 1]))
    n1, n2 = orng.box_muller(w[:, 1], w[:, 2])      # cauchy : 1 uniform + 2 normals
    return (orng.u01(w[:, 0]), n1, n2)


# ----------------------------------------------------------------------------------- the algorithm
def _costs(problem_eval, optimum, x32, noise_draws):
    """rlepso_optimizer.py:68-74 with the float32 rounding contract.  Returns float32 [NP]."""
    f = problem_eval(x32.astype(np.float64), noise_draws)
    if optimum is not None:
        f = f - optimum
    return np.asarray(f, dtype=np.float64).astype(np.float32)


def init_population(cfg: Config, problem_eval, optimum, lb, ub, tape) -> State:
    """rlepso_optimizer.py:39-65.  `problem_eval(x64, noise_draws) -> f64[NP]`."""
    NP, D = cfg.NP, cfg.dim
    pos = (lb + (ub - lb) * tape.init_pos_u()).astype(np.float32)
    vmax = 0.1 * (ub - lb)
    vel = (-vmax + (vmax - (-vmax)) * tape.init_vel_u()).astype(np.float32)
    c = _costs(problem_eval, optimum, pos, tape.noise(0))
    gi = int(np.argmin(c))
    return State(x=pos.copy(), v=vel.copy(), pbest_x=pos.copy(), c_cost=c.copy(), pbest=c.copy(),
                 pni=np.zeros(NP, dtype=np.int32), gbest_x=pos[gi].copy(), gbest_val=c[gi], gbest_idx=gi,
                 no_improve=0, fes=NP, log_index=1, curve=[c[gi]])


def coefficients(action, pni, NP):
    """rlepso_optimizer.py:112-132 under numpy-2 promotion: `action` is float32, python scalars do not
    up-cast it, so w / scale / c1..c4 are float32 values stored into float64 arrays; c_mutation is
    float32(a0*0.01) * per_no_improve(float64).  Windows are stride-5, width-7 (overlapping)."""
    action = np.asarray(action, dtype=np.float32)
    assert action.shape[-1] == N_GROUP * 7, 'actions size is not right!'
    w = np.zeros(NP); cm = np.zeros(NP)
    c1 = np.zeros(NP); c2 = np.zeros(NP); c3 = np.zeros(NP); c4 = np.zeros(NP)
    per = NP // N_GROUP
    pni64 = pni.astype(np.float64)
    for g in range(N_GROUP):
        a = action[g * N_GROUP: g * N_GROUP + 7]
        sl = slice(g * per, (g + 1) * per)
        cm[sl] = a[0] * 0.01 * pni64[sl]
        w[sl] = a[1] * 0.8 + 0.1
        scale = 1. / (a[3] + a[4] + a[5] + a[6] + 1e-5) * a[2] * 8
        c1[sl] = scale * a[3]; c2[sl] = scale * a[4]; c3[sl] = scale * a[5]; c4[sl] = scale * a[6]
    return w, cm, c1, c2, c3, c4


def fdr_targets(pbest_x64, pbest64):
    """rlepso_optimizer.py:97-103: argmin_j (pf[j]-pf[i]) / (|px[j,d]-px[i,d]| + 1e-5), first min."""
    dist = np.abs(pbest_x64[None, :, :] - pbest_x64[:, None, :])           # [i,j,d]
    delta = pbest64[None, :] - pbest64[:, None]                            # [i,j]
    return np.argmin(delta[:, :, None] / (dist + 1e-5), axis=1)            # [i,d]


def update(cfg: Config, S: State, action, problem_eval, optimum, lb, ub, tape, aux=None):
    """rlepso_optimizer.py:173-263.  Returns (new_state, next_state_feature[1], reward, done).
    `aux` (dict) receives intermediates (targets, masks) for index-parity tests."""
    NP, D = cfg.NP, cfg.dim
    pci = pci_table(NP)
    vmax = 0.1 * (ub - lb)
    x = S.x.astype(np.float64); v = S.v.astype(np.float64); px = S.pbest_x.astype(np.float64)
    pf = S.pbest.astype(np.float64); gx = S.gbest_x.astype(np.float64)
    pre_gbest = S.gbest_val

    r1 = tape.rand1()[:, None]
    r2 = tape.rand2()[:, None]
    # CLPSO  :76-95
    ru = tape.clpso_u()
    keep_own = ru > pci[:, None]
    tr = tape.tour()
    cand_cost = S.pbest[tr]                                                # float32 compare == float64 compare
    pick = np.argmin(cand_cost, axis=-1)
    tsel = np.take_along_axis(tr, pick[..., None], axis=-1)[..., 0]        # [NP,D] chosen particle
    tpos = px[tsel, np.arange(D)[None, :]]
    v_clpso = ru * (np.where(keep_own, px, tpos) - x)
    # FDR    :97-109
    tgt = fdr_targets(px, pf)
    v_fdr = tape.fdr_u() * (px[tgt, np.arange(D)[None, :]] - px)
    v_pbest = r1 * (px - x)
    v_gbest = r2 * (gx[None, :] - x)
    w, cm, c1, c2, c3, c4 = coefficients(action, S.pni, NP)
    nv = w[:, None] * v + c1[:, None] * v_clpso + c2[:, None] * v_fdr + c3[:, None] * v_gbest + c4[:, None] * v_pbest
    nv = np.clip(nv, -vmax, vmax)
    nx = np.clip(x + nv, lb, ub)
    nx32 = nx.astype(np.float32); nv32 = nv.astype(np.float32)

    fes = S.fes + NP
    nc = _costs(problem_eval, optimum, nx32, tape.noise(0))
    better = nc < S.pbest
    bi = int(np.argmin(nc)); bv = nc[bi]
    gb_better = bool(bv < S.gbest_val)
    N = State(x=nx32, v=nv32, pbest_x=np.where(better[:, None], nx32, S.pbest_x), c_cost=nc,
              pbest=np.where(better, nc, S.pbest), pni=None,
              gbest_x=nx32[bi].copy() if gb_better else S.gbest_x.copy(),
              gbest_val=bv if gb_better else S.gbest_val, gbest_idx=bi if gb_better else S.gbest_idx,
              no_improve=0 if gb_better else S.no_improve + 1, fes=fes, log_index=S.log_index,
              curve=list(S.curve))
    # patience :229-233
    improved = nc < S.c_cost
    pni = S.pni + 1
    pni = np.where(improved, 0, pni).astype(np.int32)
    N.pni = pni
    # stochastic re-initialisation :238, :134-168
    mask = tape.reinit_u() < cm * 0.01 * pni.astype(np.float64)
    if aux is not None:
        aux.update(fdr_target=tgt, tour_sel=tsel, keep_own=keep_own, better=better, reinit_mask=mask,
                   cost_main=nc)
    if np.any(mask):
        rp = (lb + (ub - lb) * tape.reinit_pos_u()).astype(np.float32)
        rv = (-vmax + (vmax - (-vmax)) * tape.reinit_vel_u()).astype(np.float32)
        px2 = np.where(mask[:, None], rp, N.x)
        pv2 = np.where(mask[:, None], rv, N.v)
        c2_ = _costs(problem_eval, optimum, px2, tape.noise(1))      # all NP re-evaluated (:140)
        N.fes = fes + int(np.sum(mask))                              # but only sum(mask) charged (:141)
        b2 = c2_ < N.pbest
        bi2 = int(np.argmin(c2_)); bv2 = c2_[bi2]
        g2 = bool(bv2 < N.gbest_val)
        N.pbest_x = np.where(b2[:, None], px2, N.pbest_x)
        N.pbest = np.where(b2, c2_, N.pbest)
        N.x, N.v, N.c_cost = px2, pv2, c2_
        if g2:
            N.gbest_val, N.gbest_x, N.gbest_idx = bv2, px2[bi2].copy(), bi2
        if aux is not None:
            aux.update(cost_reinit=c2_, better2=b2)
    # logging :241-261
    if N.fes >= N.log_index * cfg.log_interval:
        N.log_index += 1
        N.curve.append(N.gbest_val)
    if optimum is None:
        done = N.fes >= cfg.maxFEs
    else:
        done = bool(N.fes >= cfg.maxFEs or N.gbest_val <= 1e-8)
    reward = 1 if N.gbest_val < pre_gbest else -1
    if done:
        if len(N.curve) >= cfg.n_logpoint + 1:
            N.curve[-1] = N.gbest_val
        else:
            N.curve.append(N.gbest_val)
    N.done = done
    return N, np.array([N.fes / cfg.maxFEs]), reward, done
