"""Oracle: DE-DDQN backbone (DE family, ONE individual per step).  TEST INFRASTRUCTURE ONLY.

numpy restatement of `src/optimizer/de_ddqn_optimizer.py:46-220` (+ `operators/mutate.py:36-39,53-56,
124-127,178-181`, `operators/crossover.py:6-18`, `operators/boundary_control.py:5-9`) under the engine's
storage contract: X / X_gbest / X_prebest / costs are float32, arithmetic is float64 on the up-cast
state in the reference's operation order, the trial is rounded to float32 before evaluation and the cost
`eval(x) - optimum` is rounded to float32.

Reference quirks reproduced on purpose:
  * `c_prebest` is never refreshed (`self.__c_prebest = self.__c_prebest`, :135);
  * `X_prebest` starts as a *view* of row argmin(cost) of X (:61) and stays one -- so it follows later
    in-place writes to that row (:194) -- until the first generation boundary at which `X_gbest` has been
    re-bound to a trial array (:133, :198).  (`X_gbest`'s own initial view (:58) is unobservable: the row
    can only be overwritten by a trial that also re-binds `X_gbest` to the same values.)
  * the 5 indices `r` used by the mutation are the ones drawn in the *previous* `__get_state` (:85), with no
    distinctness / != pointer rejection;
  * binomial crossover with Cr = 1.0 (:12) always returns the donor, but still consumes
    `randint(dim, size=1)` and `rand(1, dim)` from np.random (crossover.py:12-15).

The operator-credit history (deques `OM`, `N_succ`, `N_tot` with maxlen gen_max = 10, window list `OM_W`
of W = 50) is kept in fixed arrays: per (op, metric, generation slot) count / sum / max of the recorded
improvements -- everything `__get_state` reads; float64 sums are accumulated in append order.

Pinned by `tests/golden/deddqn.npz` (the real `DE_DDQN_Optimizer` run on F32Problem proxies).
"""
from __future__ import annotations

import dataclasses

import numpy as np

from . import rng as orng

F_ = 0.5           # de_ddqn_optimizer.py:11
GEN_MAX = 10       # :14
W_ = 50            # :15
N_FEAT = 99


@dataclasses.dataclass
class Config:
    dim: int
    maxFEs: int
    n_logpoint: int = 50
    log_interval: int = 0
    NP: int = 100          # :13

    def __post_init__(self):
        if not self.log_interval:
            self.log_interval = self.maxFEs // self.n_logpoint


@dataclasses.dataclass
class State:
    X: np.ndarray            # [NP,D] f32
    cost: np.ndarray         # [NP]   f32
    pointer: int
    gen: int
    stagcount: int
    X_gbest: np.ndarray      # [D] f32
    c_gbest: np.float32
    c_gworst: np.float32
    prebest_row: int         # >= 0: X_prebest is still a view of this row of X ; -1: own values below
    X_prebest: np.ndarray    # [D] f32 (ignored while prebest_row >= 0)
    c_prebest: np.float32
    gbest_rebound: bool      # X_gbest has been re-bound to a trial array at least once
    r: np.ndarray            # [5] int   indices for the NEXT mutation
    n_tot: np.ndarray        # [4,GEN_MAX] i32     slot 0 = current generation
    n_succ: np.ndarray       # [4,4,GEN_MAX] i32
    om_sum: np.ndarray       # [4,4,GEN_MAX] f64
    om_max: np.ndarray       # [4,4,GEN_MAX] f64
    win: np.ndarray          # [W,6] f64  rows (action, om0..om3, trial_cost) in list order
    win_len: int
    fes: int
    log_index: int
    curve: list
    done: bool = False

    def prebest(self):
        return self.X[self.prebest_row] if self.prebest_row >= 0 else self.X_prebest

    def copy(self):
        kw = {f.name: (getattr(self, f.name).copy() if isinstance(getattr(self, f.name), np.ndarray) else getattr(self, f.name))
              for f in dataclasses.fields(self)}
        kw['curve'] = list(self.curve)
        return State(**kw)


# ----------------------------------------------------------------------------------- tapes
class NumpyTape:
    """Reference draw order on the global np.random."""

    def __init__(self, NP, D, noise_counts=(0, 0)):
        self.NP, self.D, self.nc = NP, D, noise_counts
        self.rec = {}

    def init_u(self):
        self.rec['init_u'] = np.random.rand(self.NP, self.D)                      # :48
        return self.rec['init_u']

    def crossover_draws(self):
        np.random.randint(self.D, size=1)                                         # crossover.py:12 (j_rand, unused: Cr = 1)
        np.random.rand(1, self.D)                                                 # crossover.py:15

    def r5(self):
        self.rec['r'] = np.random.randint(0, self.NP, 5)                          # :85
        return self.rec['r']

    def noise(self, n):
        nu, nn = self.nc
        d = tuple(np.random.rand(n) for _ in range(nu)) + tuple(np.random.randn(n) for _ in range(nn))
        self.rec['noise'] = d
        return d if d else None


class ArrayTape:
    def __init__(self, rec):
        self.rec = rec

    def init_u(self): return self.rec['init_u']
    def crossover_draws(self): pass
    def r5(self): return self.rec['r']
    def noise(self, n): return self.rec.get('noise') or None


class PhiloxTape:
    """CUDA production mapping (mbx_deddqn.cu): TAG_INIT per (i,d) at gen GEN_INIT -> init u ;
    TAG_DDQN idx 0,1 at step `gen` -> r[0..3], r[4] ; TAG_NOISE idx 0 (idx i at init) -> noise draws."""

    def __init__(self, NP, D, seed, lane, gen, noise_kind=0):
        self.NP, self.D = NP, D
        self.p = orng.PhiloxDraws(seed, lane, gen)
        self.noise_kind = noise_kind

    def init_u(self):
        return orng.u01(self.p.words(orng.TAG_INIT, self.NP * self.D)[:, 0]).reshape(self.NP, self.D)

    def crossover_draws(self): pass

    def r5(self):
        w = self.p.words(orng.TAG_DDQN, 2)
        return orng.bounded(np.concatenate([w[0], w[1, :1]]), self.NP)

    def noise(self, n):
        from .rlepso import philox_noise
        return philox_noise(self.p, self.noise_kind, n, orng.TAG_NOISE)


# ----------------------------------------------------------------------------------- helpers
def _cost(problem_eval, optimum, x32, draws):
    f = problem_eval(np.atleast_2d(x32).astype(np.float64), draws)
    if optimum is not None:
        f = f - optimum
    return np.asarray(f, dtype=np.float64).astype(np.float32)


def get_state(cfg: Config, S: State, lb, ub, tape):
    """de_ddqn_optimizer.py:76-129 -> float64 [99]; draws the next 5 indices into S.r."""
    NP, D = cfg.NP, cfg.dim
    X = S.X.astype(np.float64); cost = S.cost.astype(np.float64)
    gb, gw, pb = np.float64(S.c_gbest), np.float64(S.c_gworst), np.float64(S.c_prebest)
    p = S.pointer
    max_dist = np.linalg.norm(np.array([ub - lb]).repeat(D), 2)
    f = np.zeros(N_FEAT)
    with np.errstate(divide='ignore', invalid='ignore'):
        f[0] = (cost[p] - gb) / (gw - gb)
        f[1] = (np.mean(cost) - gb) / (gw - gb)
        f[2] = np.std(cost) / ((gw - gb) / 2)
        f[3] = (cfg.maxFEs - S.fes) / cfg.maxFEs
        f[4] = D / D
        f[5] = S.stagcount / cfg.maxFEs
        S.r = np.asarray(tape.r5()).astype(np.int64)
        for j in range(5):
            f[6 + j] = np.linalg.norm(X[p] - X[S.r[j]], 2) / max_dist
        f[11] = np.linalg.norm(X[p] - S.prebest().astype(np.float64), 2) / max_dist
        for j in range(5):
            f[12 + j] = (cost[p] - cost[S.r[j]]) / (gw - gb)
        f[17] = (cost[p] - pb) / (gw - gb)
        f[18] = np.linalg.norm(X[p] - S.X_gbest.astype(np.float64), 2) / max_dist
        G = min(GEN_MAX, S.gen)
        i = 19
        for op in range(4):
            for m in range(4):
                for g in range(G):
                    if S.n_tot[op, g] > 0:
                        f[i] += S.n_succ[op, m, g] / S.n_tot[op, g]
                i += 1
        for op in range(4):
            tot = int(S.n_tot[op, :G].sum())
            for m in range(4):
                for g in range(G):
                    if S.n_succ[op, m, g] > 0:
                        f[i] += S.om_sum[op, m, g]        # == sequential sum of the recorded values of slot g
                if tot > 0:
                    f[i] = f[i] / tot
                i += 1
        if S.gen >= 2:
            for op in range(4):
                for m in range(4):
                    dn = int(S.n_tot[op, 0]) - int(S.n_tot[op, 1])
                    if dn != 0 and S.n_succ[op, m, 0] > 0 and S.n_succ[op, m, 1] > 0:
                        f[i] = (S.om_max[op, m, 0] - S.om_max[op, m, 1]) / (S.om_max[op, m, 1] * np.abs(dn))
                    i += 1
        else:
            i = 67
        for op in range(4):
            for m in range(4):
                for g in range(G):
                    if S.n_succ[op, m, g] > 0:
                        f[i] += S.om_max[op, m, g]
                i += 1
        for w in range(min(W_, S.win_len)):
            a = int(S.win[w, 0])
            for m in range(4):
                f[i + a * 4 + m] += S.win[w, m + 1]
    return f


def init_population(cfg: Config, problem_eval, optimum, lb, ub, tape):
    """de_ddqn_optimizer.py:46-74 -> (State, features[99])."""
    NP, D = cfg.NP, cfg.dim
    X = (tape.init_u() * (ub - lb) + lb).astype(np.float32)
    cost = _cost(problem_eval, optimum, X, tape.noise(NP))
    gi = int(np.argmin(cost))
    S = State(X=X, cost=cost, pointer=0, gen=0, stagcount=0, X_gbest=X[gi].copy(), c_gbest=cost[gi],
              c_gworst=np.max(cost), prebest_row=gi, X_prebest=X[gi].copy(), c_prebest=cost[gi], gbest_rebound=False,
              r=np.zeros(5, dtype=np.int64), n_tot=np.zeros((4, GEN_MAX), dtype=np.int32),
              n_succ=np.zeros((4, 4, GEN_MAX), dtype=np.int32), om_sum=np.zeros((4, 4, GEN_MAX)),
              om_max=np.zeros((4, 4, GEN_MAX)), win=np.zeros((W_, 6)), win_len=0, fes=NP, log_index=1, curve=[cost[gi]])
    return S, get_state(cfg, S, lb, ub, tape)


def mutate(X, X_gbest, action, p, r):
    """mutate.py:36-39 (rand/1), :53-56 (rand/2), :124-127 (rand-to-best/2), :178-181 (cur-to-rand/1); F = 0.5."""
    x = X
    if action == 0:
        return x[r[0]] + F_ * (x[r[1]] - x[r[2]])
    if action == 1:
        return x[r[0]] + F_ * (x[r[1]] - x[r[2]] + x[r[3]] - x[r[4]])
    if action == 2:
        return x[r[0]] + F_ * (X_gbest - x[r[0]] + x[r[1]] - x[r[2]] + x[r[3]] - x[r[4]])
    if action == 3:
        return x[p] + F_ * (x[r[0]] - x[p] + x[r[1]] - x[r[2]])
    raise ValueError('Action error')                                               # :153


def update(cfg: Config, S: State, action, problem_eval, optimum, lb, ub, tape, aux=None):
    """de_ddqn_optimizer.py:131-220.  Mutates and returns S -> (S, features[99], reward, done)."""
    NP, D = cfg.NP, cfg.dim
    action = int(action)
    if S.pointer == 0:                                                             # :132-142
        if S.gbest_rebound:
            if S.prebest_row >= 0:
                S.prebest_row = -1
            S.X_prebest = S.X_gbest.copy()
        S.gen += 1
        for a in (S.n_tot, S.n_succ, S.om_sum, S.om_max):                          # appendleft on maxlen-10 deques
            a[..., 1:] = a[..., :-1].copy()
            a[..., 0] = 0
    p = S.pointer
    X64 = S.X.astype(np.float64)
    donor = mutate(X64, S.X_gbest.astype(np.float64), action, p, S.r)
    donor = np.clip(donor, lb, ub)                                                 # :155
    tape.crossover_draws()
    trial32 = donor.astype(np.float32)                                             # Cr = 1: trial == donor (:157)
    tc = _cost(problem_eval, optimum, trial32, tape.noise(1))[0]
    S.fes += 1
    cp = S.cost[p]
    reward = max(np.float64(cp) - np.float64(tc), 0)                               # :165
    S.n_tot[action, 0] += 1
    om = np.array([np.float64(cp) - tc, np.float64(S.c_prebest) - tc, np.float64(S.c_gbest) - tc,
                   np.median(S.cost.astype(np.float64)) - tc], dtype=np.float64)
    for m in range(4):
        if om[m] > 0:
            if S.n_succ[action, m, 0] == 0:
                S.om_sum[action, m, 0] = om[m]; S.om_max[action, m, 0] = om[m]
            else:
                S.om_sum[action, m, 0] += om[m]; S.om_max[action, m, 0] = max(S.om_max[action, m, 0], om[m])
            S.n_succ[action, m, 0] += 1
    if S.win_len >= W_:                                                            # :176-185
        same = np.nonzero(S.win[:S.win_len, 0] == action)[0]
        k = int(same[0]) if same.size else int(np.argmax(S.win[:S.win_len, 5]))
        S.win[k:S.win_len - 1] = S.win[k + 1:S.win_len].copy()
        S.win_len -= 1
    S.win[S.win_len] = [action, om[0], om[1], om[2], om[3], np.float64(tc)]
    S.win_len += 1
    if tc >= S.c_gbest:                                                            # :188-189
        S.stagcount += 1
    replaced = bool(tc <= cp)
    if replaced:                                                                   # :192-198
        S.cost[p] = tc
        S.X[p] = trial32
        if tc <= S.c_gbest:
            S.c_gbest = tc
            S.X_gbest = trial32.copy()
            S.gbest_rebound = True
    if tc > S.c_gworst:                                                            # :200-201
        S.c_gworst = tc
    S.pointer = (S.pointer + 1) % NP
    if S.fes >= S.log_index * cfg.log_interval:
        S.log_index += 1
        S.curve.append(S.c_gbest)
    if optimum is None:
        done = S.fes >= cfg.maxFEs
    else:
        done = bool(S.fes >= cfg.maxFEs or S.c_gbest <= 1e-8)
    feat = get_state(cfg, S, lb, ub, tape)
    if done:
        if len(S.curve) >= cfg.n_logpoint + 1:
            S.curve[-1] = S.c_gbest
        else:
            S.curve.append(S.c_gbest)
    S.done = done
    if aux is not None:
        aux.update(trial=trial32, trial_cost=tc, replaced=replaced, om=om)
    return S, feat, reward, done
