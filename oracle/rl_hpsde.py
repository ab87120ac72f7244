"""Oracle: RL_HPSDE backbone (DE family, whole population per step, success-history memories, linear population-size reduction,
landscape state from a progressive random walk).  TEST INFRASTRUCTURE ONLY.

numpy restatement of `src/optimizer/rl_hpsde_optimizer.py:7-287` (+ `operators/mutate.py:12-32,148-157,184-188`,
`operators/crossover.py:6-18`, `operators/boundary_control.py:5-9`; the two state statistics are `oracle/landscape.py` `dfdc_r` / `drie_r`)
under the engine's storage contract: population / costs float32, donor arithmetic float64 in the reference's operation order, everything
that is evaluated (trial vectors, walk points) is rounded to float32 first and kept in that rounded form.

What the reference does, as restated here:
  * NP starts at 18 * dim and shrinks linearly to 4 with the consumed budget (`LPSR`, :90-99); the population is re-sorted by cost every
    generation; memories MF / MCr of H = dim / 2 slots start at 0.5, slot k takes the weighted Lehmer means of the successful F / Cr of a
    generation (k advances) or is reset to 0.5 when nothing improved (k stays) (:76-87);
  * per individual: slot r ~ U{0..H-1}; Cr ~ clip(N(MCr[r], 0.1), 0, 1); F ~ Cauchy(MF[r], 0.1) (actions 0, 1) or Levy(MF[r], 0.1) (actions
    2, 3), a negative F is reflected about its location once, then min(F, 1) (:52-68);
  * actions 0 / 2: current-to-rand/1, actions 1 / 3: current-to-best/1 with 3 / 2 mutually distinct indices != i (`generate_random_int`);
  * selection is strict (`ncost < cost`), reward = fraction of improved individuals;
  * the state after every generation: a progressive random walk of 201 points through the WHOLE box (:150-167) is evaluated (201 charged
    evaluations), state = [DFDC r > 0.15] + 2 * [DRIE entropy >= 0.5] (:169-229).
`np.argsort` ties (equal costs) are ordered by index here (stable); the reference's order is unspecified.

Pinned by `tests/golden/rl_hpsde.npz` (the real `RL_HPSDE_Optimizer` on F32Problem proxies, tests/golden/make_golden_hpsde.py).
"""
from __future__ import annotations

import dataclasses

import numpy as np

from . import landscape as ola

RW_STEPS, STEP_SIZE = 200, 10          # rl_hpsde_optimizer.py:110-111
NP_MIN = 4                             # :108


@dataclasses.dataclass
class Config:
    dim: int
    maxFEs: int
    n_logpoint: int = 50
    log_interval: int = 0

    def __post_init__(self):
        if not self.log_interval:
            self.log_interval = self.maxFEs // self.n_logpoint
        self.NP_max = 18 * self.dim                    # :107
        self.H = int(self.dim * 0.5)                   # :18-19 (Hm = 0.5)


@dataclasses.dataclass
class State:
    group: np.ndarray        # [NP,D] f32, sorted by cost
    cost: np.ndarray         # [NP]   f32
    MF: np.ndarray           # [H] f64
    MCr: np.ndarray          # [H] f64
    k: int
    gbest: np.float32
    fes: int
    log_index: int
    curve: list
    done: bool = False

    @property
    def NP(self):
        return len(self.cost)

    def copy(self):
        kw = {f.name: (getattr(self, f.name).copy() if isinstance(getattr(self, f.name), np.ndarray) else getattr(self, f.name))
              for f in dataclasses.fields(self)}
        kw['curve'] = list(self.curve)
        return State(**kw)


def f32(a):
    return np.asarray(a, dtype=np.float64).astype(np.float32)


class NumpyTape:
    """Reference draw order on the global np.random (scipy's cauchy / levy `rvs` draw from it too); `rec` keeps the derived quantities
    (F, Cr, index matrix, jrand, crossover uniforms, walk draws) so that the batched engine can replay them."""

    def __init__(self, D, noise_counts=(0, 0)):
        self.D, self.nc = D, noise_counts
        self.rec = {}

    def init_u(self, NP):
        self.rec['init_u'] = np.random.rand(NP, self.D)
        return self.rec['init_u']

    def f_cr(self, MF, MCr, NP, dist):
        import scipy.stats as stats
        ind_r = np.random.randint(0, MF.shape[0], size=NP)                                          # :55
        cr_raw = np.random.normal(loc=MCr[ind_r], scale=0.1, size=NP)                               # :56
        locs = MF[ind_r]
        f_raw = stats.cauchy.rvs(loc=locs, scale=0.1, size=NP) if dist == 'cauchy' else stats.levy.rvs(loc=locs, scale=0.1, size=NP)
        self.rec.update(ind_r=ind_r, cr_raw=cr_raw, f_raw=f_raw)
        return ind_r, cr_raw, f_raw

    def r(self, NP, cols):
        r = np.random.randint(low=0, high=NP, size=(NP, cols))                                      # mutate.py:21-31
        for col in range(cols):
            while True:
                rep = [np.equal(r[:, col], r[:, i]) for i in range(col)]
                rep.append(np.equal(r[:, col], np.arange(NP)))
                idx = np.nonzero(np.any(np.stack(rep), axis=0))[0]
                if idx.size:
                    r[idx, col] = np.random.randint(low=0, high=NP, size=idx.size)
                else:
                    break
        self.rec['r'] = np.concatenate([r, np.zeros((NP, 3 - cols), dtype=r.dtype)], axis=1)
        return r

    def crossover(self, NP):
        self.rec['jrand'] = np.random.randint(self.D, size=NP)                                      # crossover.py:11
        self.rec['cr_u'] = np.random.rand(NP, self.D)                                               # crossover.py:14
        return self.rec['jrand'], self.rec['cr_u']

    def walk(self):
        zone_u = np.random.rand(self.D)                                                             # :152
        r_u = np.random.rand(self.D)                                                                # :155
        rD = int(np.random.choice(self.D, 1)[0])                                                    # :157
        step_u = np.stack([np.random.rand(self.D) for _ in range(RW_STEPS)])                        # :163
        self.rec.update(zone_u=zone_u, r_u=r_u, rD=rD, step_u=step_u)
        return zone_u, r_u, rD, step_u

    def noise(self, n, key):
        nu, nn = self.nc
        d = tuple(np.random.rand(n) for _ in range(nu)) + tuple(np.random.randn(n) for _ in range(nn))
        self.rec[key] = d
        return d if d else None


class ArrayTape:
    def __init__(self, rec):
        self.rec = rec

    def init_u(self, NP): return np.asarray(self.rec['init_u'], dtype=np.float64)
    def f_cr(self, MF, MCr, NP, dist): return np.asarray(self.rec['ind_r']), np.asarray(self.rec['cr_raw']), np.asarray(self.rec['f_raw'])
    def r(self, NP, cols): return np.asarray(self.rec['r'])[:, :cols]
    def crossover(self, NP): return np.asarray(self.rec['jrand']), np.asarray(self.rec['cr_u'], dtype=np.float64)
    def walk(self): return (np.asarray(self.rec['zone_u']), np.asarray(self.rec['r_u']), int(self.rec['rD']), np.asarray(self.rec['step_u']))
    def noise(self, n, key): return self.rec.get(key) or None


def progressive_walk(lb, ub, zone_u, r_u, rD, step_u):
    """:150-167 with the draws given -> [RW_STEPS + 1, D] float64 walk (before the float32 rounding of the evaluation)."""
    D = len(zone_u)
    zone = np.where(zone_u < 0.5, -1.0, 1.0)
    s = np.zeros((RW_STEPS + 1, D))
    s[0] = (ub + lb) / 2 + zone * (r_u * (ub - lb) / 2)
    s[0][rD] = lb if zone[rD] == -1 else ub
    for k in range(1, RW_STEPS + 1):
        s[k] = s[k - 1] + step_u[k - 1] * (-STEP_SIZE) * zone
        hi, lo = s[k] > ub, s[k] < lb
        s[k][hi] = 2 * ub - s[k][hi]
        s[k][lo] = 2 * lb - s[k][lo]
        zone[hi | lo] *= -1
    return s


def get_state(cfg, S, ev, optimum, lb, ub, tape, key):
    """:218-229 -> (state in 0..3, dfdc r, drie entropy); charges RW_STEPS + 1 evaluations (done by the caller)."""
    sample = f32(progressive_walk(lb, ub, *tape.walk())).astype(np.float64)          # the evaluated (float32) points are the walk
    f = ev(sample, tape.noise(RW_STEPS + 1, key))
    c = f32(f if optimum is None else f - optimum).astype(np.float64)
    best = S.group[int(np.argmin(S.cost))].astype(np.float64)
    r = ola.dfdc_r(sample, c, best)
    h = ola.drie_r(c)
    if not (-1 <= r <= 1):
        raise ValueError(f'DFDC error: {r}')                                          # :180 (degenerate walk: zero variance)
    return int(0.15 < r) * 1 + int(h >= 0.5) * 2, r, h


def sort_by_cost(S):
    ind = np.argsort(S.cost, kind='stable')
    S.group, S.cost = S.group[ind], S.cost[ind]
    S.gbest = np.float32(S.cost.min())


def init_population(cfg: Config, ev, optimum, lb, ub, tape):
    """:126-134.  `ev(x64 [n,D], noise_draws) -> f64 [n]` evaluates float32-representable positions."""
    NP = cfg.NP_max
    group = f32(tape.init_u(NP) * (ub - lb) + lb)
    f = ev(group.astype(np.float64), tape.noise(NP, 'noise_init'))
    cost = f32(f if optimum is None else f - optimum)
    S = State(group=group, cost=cost, MF=np.ones(cfg.H) * 0.5, MCr=np.ones(cfg.H) * 0.5, k=0, gbest=np.float32(cost.min()), fes=NP,
              log_index=1, curve=[np.float32(cost.min())])
    sort_by_cost(S)
    state, r, h = get_state(cfg, S, ev, optimum, lb, ub, tape, 'noise_walk')
    S.fes += RW_STEPS + 1
    return S, state


def mean_wL(df, s):
    w = df / np.sum(df)
    return np.sum(w * (s ** 2)) / np.sum(w * s) if np.sum(w * s) > 0.000001 else 0.5                # :70-74


def update(cfg: Config, S: State, action, ev, optimum, lb, ub, tape):
    """:231-287 -> (new State, state in 0..3, reward, done)."""
    if action not in (0, 1, 2, 3):
        raise ValueError(f'action error: {action}')
    S = S.copy()
    NP = S.NP
    x = S.group.astype(np.float64)
    ind_r, cr_raw, f_raw = tape.f_cr(S.MF, S.MCr, NP, 'cauchy' if action < 2 else 'levy')
    Cr = np.minimum(1, np.maximum(0, cr_raw))
    locs = S.MF[ind_r]
    F = np.array(f_raw, dtype=np.float64)
    neg = F < 0
    F[neg] = 2 * locs[neg] - F[neg]
    F = np.minimum(1, F)
    Fc = F.reshape(-1, 1)
    if action in (0, 2):
        r = tape.r(NP, 3)
        v = x + Fc * (x[r[:, 0]] - x + x[r[:, 1]] - x[r[:, 2]])                                     # mutate.py:184-188
    else:
        r = tape.r(NP, 2)
        best = x[int(np.argmin(S.cost))]
        v = x + Fc * (best - x + x[r[:, 0]] - x[r[:, 1]])                                           # mutate.py:148-157
    v = np.clip(v, lb, ub)
    jrand, cr_u = tape.crossover(NP)
    u = np.where(cr_u < Cr.reshape(-1, 1), v, x)
    u[np.arange(NP), jrand] = v[np.arange(NP), jrand]
    u32 = f32(u)
    f = ev(u32.astype(np.float64), tape.noise(NP, 'noise_trial'))
    ncost = f32(f if optimum is None else f - optimum)
    S.fes += NP
    optim = np.where(ncost < S.cost)[0]
    SF, SCr = F[optim], Cr[optim]
    df = np.maximum(0, S.cost.astype(np.float64) - ncost.astype(np.float64))[optim]
    if SF.shape[0] > 0:                                                                             # :76-87
        S.MF[S.k] = mean_wL(df, SF)
        S.MCr[S.k] = mean_wL(df, SCr)
        S.k = (S.k + 1) % S.MF.shape[0]
    else:
        S.MF[S.k] = 0.5
        S.MCr[S.k] = 0.5
    S.group[optim] = u32[optim]
    S.cost = np.minimum(S.cost, ncost)
    sort_by_cost(S)                                                                                 # LPSR (:90-99)
    N = max(int(cfg.NP_max + np.round(NP_MIN - cfg.NP_max) * S.fes / cfg.maxFEs), 1)
    if N < NP:
        S.group, S.cost = S.group[:N].copy(), S.cost[:N].copy()
    if S.fes >= S.log_index * cfg.log_interval:
        S.log_index += 1
        S.curve.append(S.gbest)
    reward = optim.shape[0] / NP
    S.done = bool(S.fes >= cfg.maxFEs) if optimum is None else bool(S.fes >= cfg.maxFEs or S.gbest <= 1e-8)
    if S.done:
        if len(S.curve) >= cfg.n_logpoint + 1:
            S.curve[-1] = S.gbest
        else:
            S.curve.append(S.gbest)
    state, r, h = get_state(cfg, S, ev, optimum, lb, ub, tape, 'noise_walk')
    S.fes += RW_STEPS + 1
    return S, state, reward, S.done
