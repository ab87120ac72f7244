"""Oracle: NRLPSO backbone (PSO family, ONE particle per step, neighbourhood-based learning targets, Q-learning over 4 strategies).
TEST INFRASTRUCTURE ONLY -- groundwork for a CUDA path (none exists yet: DESIGN.md section 10).

numpy restatement of `src/optimizer/nrlpso_optimizer.py:7-297` under the engine's storage contract (positions / velocities / costs float32,
arithmetic float64 on the up-cast state in the reference's operation order, every evaluated point rounded to float32 first).

Reference behaviour reproduced on purpose:
  * `__gbest_pos` is a VIEW of a population row whenever it was bound by `self.__population[...]` (:46, :271): it follows every later
    in-place write to that row (the particle's own move :243-244, a neighbourhood mutant that replaces the row :227, :206) until it is
    re-bound; `neb_mutation` re-binds it to a fresh array (:222).  Kept here as `gbest_row >= 0`;
  * `__pbest_cost` is NOT refreshed when a particle improves on its personal best (:259-263 copy the position only); only
    `neb_mutation` writes it (:203);
  * the neighbourhood tables (`pbest_neb`, `gbest_neb`: the k = 5 nearest CURRENT positions to every pbest / to gbest) are snapshots taken
    when the pointer wraps (:239-240) and used, stale, for the next NP steps; the inertia weight follows a logistic map (:88-93);
  * the crowding term `ef` (:107-121) needs the full pairwise distance matrix before and after the move;
  * the returned state is the stored strategy of the NEXT particle (:272-273, :297); rewards 2 / 1 / 0 / -2 (:95-105);
  * every step charges 1 evaluation, a stagnating particle (2 steps without a new personal best) 2 more (:265-266, :192-230).
`np.argsort` ties are ordered by index here (stable); the reference's order is unspecified.

Pinned by `tests/golden/nrlpso.npz` (the real `NRLPSO_Optimizer` on F32Problem proxies, tests/golden/make_golden_nrlpso.py).
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

NP_, K_, N_STATE = 100, 5, 4                                   # nrlpso_optimizer.py:12-13, 20
W_MAX, W_MIN, U_, V_ = 1.0, 0.4, 0.6, 0.33                     # :15-18


@dataclasses.dataclass
class Config:
    dim: int
    maxFEs: int
    n_logpoint: int = 50
    log_interval: int = 0

    def __post_init__(self):
        if not self.log_interval:
            self.log_interval = self.maxFEs // self.n_logpoint


@dataclasses.dataclass
class State:
    pop: np.ndarray            # [NP,D] f32
    vel: np.ndarray            # [NP,D] f32
    cost: np.ndarray           # [NP]   f32
    pbest_pos: np.ndarray      # [NP,D] f32
    pbest_cost: np.ndarray     # [NP]   f32
    gbest_row: int             # >= 0: gbest_pos is a view of this population row ; -1: own values below
    gbest_pos: np.ndarray      # [D] f32 (ignored while gbest_row >= 0)
    gbest_cost: np.float32
    stag: np.ndarray           # [NP] int
    qstate: np.ndarray         # [NP] int (last strategy applied to each particle; random at the start)
    r_w: float
    w: float
    pbest_neb_index: np.ndarray    # [NP,K] int
    pbest_neb: np.ndarray          # [NP,K,D] f32 snapshot
    gbest_neb_index: np.ndarray    # [K] int
    gbest_neb: np.ndarray          # [K,D] f32 snapshot
    pointer: int
    fes: int
    log_index: int
    curve: list
    done: bool = False

    def gbest(self):
        return self.pop[self.gbest_row] if self.gbest_row >= 0 else self.gbest_pos

    def copy(self):
        kw = {f.name: (getattr(self, f.name).copy() if isinstance(getattr(self, f.name), np.ndarray) else getattr(self, f.name))
              for f in dataclasses.fields(self)}
        kw['curve'] = list(self.curve)
        return State(**kw)


def f32(a):
    return np.asarray(a, dtype=np.float64).astype(np.float32)


class NumpyTape:
    """Reference draw order on the global np.random (the oracle asks for the draws in the reference's order); `rec` keeps them by name so
    that the batched engine can replay a step."""

    def __init__(self, D, noise_counts=(0, 0)):
        self.D, self.nc = D, noise_counts
        self.rec = {}

    def init_u(self):
        self.rec['init_u'] = np.random.rand(NP_, self.D)
        return self.rec['init_u']

    def init_rw(self):
        self.rec['r_w'] = float(np.random.rand())
        return self.rec['r_w']

    def init_state(self):
        self.rec['qstate'] = np.random.randint(low=0, high=N_STATE, size=NP_)
        return self.rec['qstate']

    def scalars(self):
        self.rec['r1'], self.rec['r2'] = float(np.random.rand()), float(np.random.rand())                # :138-139
        return self.rec['r1'], self.rec['r2']

    def pick(self, name):                                                                                # get_p_a / get_p_b (:126-132)
        self.rec[name] = int(np.random.randint(0, K_))
        return self.rec[name]

    def vec(self, name):                                                                                 # r1 / r2 of strategy 3, mutation steps
        self.rec[name] = np.random.rand(self.D)
        return self.rec[name]

    def noise(self, n):
        nu, nn = self.nc
        d = tuple(np.random.rand(n) for _ in range(nu)) + tuple(np.random.randn(n) for _ in range(nn))
        return d if d else None


class ArrayTape:
    """Replays named draws (a NumpyTape's `rec`, or the tape the batched engine drew on the device)."""

    def __init__(self, rec):
        self.rec = rec

    def init_u(self): return np.asarray(self.rec['init_u'], dtype=np.float64)
    def init_rw(self): return float(self.rec['r_w'])
    def init_state(self): return np.asarray(self.rec['qstate'])
    def scalars(self): return float(self.rec['r1']), float(self.rec['r2'])
    def pick(self, name): return int(self.rec[name])
    def vec(self, name): return np.asarray(self.rec[name], dtype=np.float64)
    def noise(self, n): return None


def _cost(ev, optimum, x32, tape):
    f = ev(x32.astype(np.float64)[None, :], tape.noise(1))
    return f32(f if optimum is None else f - optimum)[0]


def init_population(cfg: Config, ev, optimum, lb, ub, tape):
    """:30-61 -> (State, state of particle 0)."""
    D = cfg.dim
    pop = f32(tape.init_u() * (ub - lb) + lb)
    f = ev(pop.astype(np.float64), tape.noise(NP_))
    cost = f32(f if optimum is None else f - optimum)
    g = int(np.argmin(cost))
    r_w = tape.init_rw()
    qstate = np.asarray(tape.init_state(), dtype=np.int64)
    z = np.zeros
    S = State(pop=pop, vel=z((NP_, D), np.float32), cost=cost, pbest_pos=pop.copy(), pbest_cost=cost.copy(), gbest_row=g, gbest_pos=pop[g].copy(),
              gbest_cost=np.float32(cost[g]), stag=z(NP_, np.int64), qstate=qstate, r_w=r_w, w=0.0,
              pbest_neb_index=z((NP_, K_), np.int64), pbest_neb=z((NP_, K_, D), np.float32), gbest_neb_index=z(K_, np.int64),
              gbest_neb=z((K_, D), np.float32), pointer=0, fes=NP_, log_index=1, curve=[np.float32(cost[g])])
    return S, int(qstate[0])


def _neighbourhoods(S):
    """:64-85"""
    x, pb = S.pop.astype(np.float64), S.pbest_pos.astype(np.float64)
    d = np.sqrt(np.sum((pb[None, :] - x[:, None]) ** 2, axis=-1))
    d[np.arange(NP_), np.arange(NP_)] = math.inf
    S.pbest_neb_index = np.argsort(d, -1, kind='stable')[:, :K_]
    S.pbest_neb = S.pop[S.pbest_neb_index].copy()
    dg = np.sqrt(np.sum((S.gbest().astype(np.float64)[None, :] - x) ** 2, axis=-1))
    S.gbest_neb_index = np.argsort(dg, -1, kind='stable')[:K_]
    S.gbest_neb = S.pop[S.gbest_neb_index].copy()


def _ef(S, i, xi=None):
    """:107-121: normalised mean distance of particle i to the rest of the swarm (`xi`: particle i's position where it differs from the
    stored float32 row -- the reference measures the moved particle before the evaluation rounds it)."""
    x = S.pop.astype(np.float64)
    if xi is not None:
        x[i] = xi
    dist = np.sum(np.sqrt(np.sum((x[None, :] - x[:, None]) ** 2, -1)), -1) / (NP_ - 1)
    return (dist[i] - np.min(dist)) / (np.max(dist) - np.min(dist))


def _cs(p1, p2):
    return np.sum(p1 * p2) / (np.sqrt(np.sum(p1 ** 2)) * np.sqrt(np.sum(p2 ** 2)))        # :123-124


def _neb_mutation(cfg, S, i, ev, optimum, tape):
    """:192-230"""
    pb = S.pbest_pos[i].astype(np.float64)
    nb = S.pbest_neb[i].astype(np.float64)
    order = np.argsort(np.sqrt(np.sum((pb[None, :] - nb) ** 2, axis=-1)), kind='stable')
    P3 = f32(pb + tape.vec('mut_p') * (nb[order[0]] - nb[order[-1]]))
    S.fes += 1
    c = _cost(ev, optimum, P3, tape)
    if c < S.pbest_cost[i]:
        S.pbest_pos[i], S.pbest_cost[i] = P3, c
    else:
        j = int(S.pbest_neb_index[i][order[-1]])
        S.pop[j], S.cost[j] = P3, c
    g = S.gbest().astype(np.float64)
    nb = S.gbest_neb.astype(np.float64)
    order = np.argsort(np.sqrt(np.sum((g[None, :] - nb) ** 2, axis=-1)), kind='stable')
    P3 = f32(g + tape.vec('mut_g') * (nb[order[0]] - nb[order[-1]]))
    S.fes += 1
    c = _cost(ev, optimum, P3, tape)
    if c < S.gbest_cost:
        S.gbest_pos, S.gbest_row, S.gbest_cost = P3.copy(), -1, np.float32(c)
    else:
        j = int(S.gbest_neb_index[order[-1]])
        S.pop[j], S.cost[j] = P3, c


def update(cfg: Config, S: State, action, ev, optimum, lb, ub, tape):
    """:232-297 -> (new State, state of the next particle, reward, done)."""
    S = S.copy()
    i, D = S.pointer, cfg.dim
    if i == 0:
        _neighbourhoods(S)
        S.r_w = 4 * S.r_w * (1 - S.r_w)                                                    # :88-93
        S.w = U_ - ((S.fes / cfg.maxFEs) * S.r_w * W_MIN + V_ * (W_MAX - W_MIN) * (S.fes / cfg.maxFEs))
    w = S.w
    v_max = 0.1 * (ub - lb)
    # ---- generate_v_vector (:136-187)
    r1, r2 = tape.scalars()
    x = S.pop[i].astype(np.float64)
    v = S.vel[i].astype(np.float64)
    pb, gb = S.pbest_pos[i].astype(np.float64), S.gbest().astype(np.float64)
    cs = _cs(pb, gb)
    p_a = lambda: S.gbest_neb[tape.pick('pick_a')].astype(np.float64)
    p_b = lambda: S.pbest_neb[i][tape.pick('pick_b')].astype(np.float64)
    if action == 0:
        c1, c2 = 2.2, 1.8
        if cs < 0:
            P1 = pb; P2 = p_a()
            v = w * v + c1 * r1 * (P1 - x) + c2 * r2 * (P2 - x)
        else:
            v = w * v + c1 * r1 * (p_b() - x)
    elif action == 1:
        c1, c2 = 2.1, 1.8
        if cs < 0:
            P1 = p_b()
            v = w * v + c1 * r1 * (P1 - x) + c2 * r2 * (gb - x)
        else:
            v = w * v + c2 * r2 * (p_a() - x)
    elif action == 2:
        c1, c2 = 2, 2
        if cs < 0:
            v = w * v + c1 * r1 * (pb - x) + c2 * r2 * (gb - x)
        else:
            v = w * v + c2 * r2 * (gb - x)
    elif action == 3:
        c1, c2 = 1.8, 2.2
        P1 = p_b(); P2 = p_a()
        r1 = tape.vec('r1v'); r2 = tape.vec('r2v')
        v = w * v + c1 * r1 * (P1 - x) + c2 * r2 * (P2 - x)
    else:
        raise ValueError(f'action error: {action}')                                        # (the reference would fail on `None` arithmetic)
    v = np.clip(v, -v_max, v_max)
    S.vel[i] = f32(v)                                                                      # stored float32; the move below uses the float64 value
    ef_old = _ef(S, i)
    x_new = np.clip(x + v, lb, ub)
    ef_new = _ef(S, i, x_new)
    S.pop[i] = f32(x_new)
    f_old = S.cost[i]
    S.fes += 1
    f_new = _cost(ev, optimum, S.pop[i], tape)
    cond1, cond2 = f_new < f_old, ef_new > ef_old
    reward = 2 if (cond1 and cond2) else 1 if cond1 else 0 if cond2 else -2               # :95-105
    S.cost[i] = f_new
    if f_new < S.pbest_cost[i]:
        S.pbest_pos[i] = S.pop[i]
        S.stag[i] = 0
    else:
        S.stag[i] += 1
    if S.stag[i] >= 2:
        _neb_mutation(cfg, S, i, ev, optimum, tape)
    if f_new < S.gbest_cost:
        S.gbest_cost, S.gbest_row = np.float32(f_new), i
    S.qstate[i] = action
    S.pointer = (i + 1) % NP_
    if S.fes >= S.log_index * cfg.log_interval:
        S.log_index += 1
        S.curve.append(S.gbest_cost)
    S.done = bool(S.fes >= cfg.maxFEs) if optimum is None else bool(S.fes >= cfg.maxFEs or S.gbest_cost <= 1e-8)
    if S.done:
        if len(S.curve) >= cfg.n_logpoint + 1:
            S.curve[-1] = S.gbest_cost
        else:
            S.curve.append(S.gbest_cost)
    return S, int(S.qstate[S.pointer]), reward, S.done
