"""Oracle: LDE backbone (DE family, whole population per step).  TEST INFRASTRUCTURE ONLY.

numpy restatement of `src/optimizer/lde_optimizer.py:132-198` (+ helpers :21-130) under the engine's
storage contract: `pop` / `fit` live in float32, every arithmetic chain runs in float64 on the up-cast
state in the reference's operation order, trial vectors are rounded to float32 before evaluation and
costs `eval(x) - optimum` are rounded to float32; selection / sorting / histograms use those float32
costs.  Random numbers come from a tape (`NumpyTape` = reference call order on the global RNGs,
`PhiloxTape` = the CUDA production mapping, `ArrayTape` = replay).

Pinned by `tests/golden/lde.npz` (the real `LDE_Optimizer`, teacher-forced; tests/test_oracle_golden.py).

Documented deviations: `np.argsort` is the default introsort in the reference --
tie order between *exactly equal* costs is implementation-defined; the oracle and the kernel use the
stable order (index ascending).  The mutation mat-mul `M @ pop` has two non-zero terms per row; the
oracle evaluates `(1-sf)*x_i + sf*x_p` with separately rounded products (a BLAS FMA kernel may differ by
1 ulp of float64, invisible after the float32 rounding of the trial except with probability ~2^-29).
"""
from __future__ import annotations

import dataclasses

import numpy as np

from . import rng as orng

BINS = 5           # lde_optimizer.py:11
P_INI = 1.0        # :12
P_NUM_MIN = 2      # :13


@dataclasses.dataclass
class Config:
    dim: int
    maxFEs: int
    n_logpoint: int = 50
    log_interval: int = 0
    NP: int = 50           # lde_optimizer.py:10 (BASELINE config C4 asks for 100: the kernel takes any NP <= 128)

    def __post_init__(self):
        if not self.log_interval:
            self.log_interval = self.maxFEs // self.n_logpoint

    @property
    def P_MIN(self):
        return P_NUM_MIN / self.NP


@dataclasses.dataclass
class State:
    pop: np.ndarray          # [NP,D] f32, sorted by fit (ascending) after every init/update
    fit: np.ndarray          # [NP]   f32
    hist_sum: np.ndarray     # [BINS] f64: column sums of past_histo (integers; first row NP/BINS each)
    hist_cnt: int            # rows in past_histo
    gbest: np.float32
    fes: int
    log_index: int
    curve: list
    done: bool = False

    def copy(self):
        return dataclasses.replace(self, pop=self.pop.copy(), fit=self.fit.copy(), hist_sum=self.hist_sum.copy(),
                                   curve=list(self.curve))


# ----------------------------------------------------------------------------------- tapes
class NumpyTape:
    """Reference draw order on the global np.random / torch RNG (one lane, BATCH_SIZE = 1)."""

    def __init__(self, NP, D, noise_counts=(0, 0)):
        self.NP, self.D, self.nc = NP, D, noise_counts
        self.rec = {}

    def _keep(self, k, a):
        self.rec[k] = a
        return a

    def init_u(self):
        return self._keep('init_u', np.random.uniform(size=(self.NP, self.D)))          # :67

    def pbest_idx(self, high):
        return self._keep('pbest_idx', np.random.randint(0, high, size=(1, self.NP))[0])   # :90

    def donors(self):
        """torch.randint + column-wise rejection (:109-123) -> final r [NP,2]."""
        import torch
        NP = self.NP
        r = torch.randint(high=NP, size=[1, NP, 2])
        pop_index = torch.arange(NP)
        for col in range(0, 2):
            while True:
                is_repeated = [torch.eq(r[:, :, col], r[:, :, i]) for i in range(col)]
                is_repeated.append(torch.eq(r[:, :, col], pop_index))
                repeated_index = torch.nonzero(torch.any(torch.stack(is_repeated), dim=0))
                n = repeated_index.size(0)
                if n != 0:
                    r[repeated_index[:, 0], repeated_index[:, 1], col] = torch.randint(high=NP, size=[n])
                else:
                    break
        return self._keep('donors', r.numpy()[0].astype(np.int64))

    def cross_u(self):
        return self._keep('cross_u', np.random.uniform(size=(1, self.NP, self.D))[0])    # :44

    def jrand(self):
        return self._keep('jrand', np.random.randint(low=0, high=self.D, size=[1, self.NP])[0])   # :47

    def noise(self, which=0):
        nu, nn = self.nc
        d = tuple(np.random.rand(self.NP) for _ in range(nu)) + tuple(np.random.randn(self.NP) for _ in range(nn))
        self.rec['noise0'] = d
        return d if d else None


class ArrayTape:
    def __init__(self, rec):
        self.rec = rec

    def init_u(self): return self.rec['init_u']
    def pbest_idx(self, high): return self.rec['pbest_idx']
    def donors(self): return self.rec['donors']
    def cross_u(self): return self.rec['cross_u']
    def jrand(self): return self.rec['jrand']
    def noise(self, which=0): return self.rec.get('noise0') or None


class PhiloxTape:
    """Counter-based draws identical to the CUDA production path (mbx_lde.cu).
    TAG_LDE_P  per i    : w0 -> p-best pick, w1 -> r1, w2 -> r2, w3 -> j_rand
    TAG_LDE_PD per (i,d): w0 -> crossover u
    Donors are drawn without rejection: r1 uniform over {j != i}, r2 uniform over {j != i, j != r1} --
    the distribution the reference's rejection loop converges to."""

    def __init__(self, NP, D, seed, lane, gen, noise_kind=0):
        self.NP, self.D = NP, D
        self.p = orng.PhiloxDraws(seed, lane, gen)
        self.noise_kind = noise_kind
        self._p = None

    def _P(self):
        if self._p is None:
            self._p = self.p.words(orng.TAG_LDE_P, self.NP)
        return self._p

    def init_u(self):
        return orng.u01(self.p.words(orng.TAG_INIT, self.NP * self.D)[:, 0]).reshape(self.NP, self.D)

    def pbest_idx(self, high):
        return orng.bounded(self._P()[:, 0], high)

    def donors(self):
        NP = self.NP
        i = np.arange(NP)
        r1 = (i + 1 + orng.bounded(self._P()[:, 1], NP - 1)) % NP
        k = orng.bounded(self._P()[:, 2], NP - 2)
        a, b = np.minimum(i, r1), np.maximum(i, r1)
        k = k + (k >= a)
        k = k + (k >= b)
        return np.stack([r1, k], axis=1).astype(np.int64)

    def cross_u(self):
        return orng.u01(self.p.words(orng.TAG_LDE_PD, self.NP * self.D)[:, 0]).reshape(self.NP, self.D)

    def jrand(self):
        return orng.bounded(self._P()[:, 3], self.D)

    def noise(self, which=0):
        from .rlepso import philox_noise
        return philox_noise(self.p, self.noise_kind, self.NP, orng.TAG_NOISE)


# ----------------------------------------------------------------------------------- helpers
def _costs(problem_eval, optimum, x32, noise_draws):
    f = problem_eval(x32.astype(np.float64), noise_draws)
    if optimum is not None:
        f = f - optimum
    return np.asarray(f, dtype=np.float64).astype(np.float32)


def order_by_f(pop, fit):
    """lde_optimizer.py:73-78 with the stable tie rule."""
    o = np.argsort(fit, kind='stable')
    return pop[o], fit[o], o


def maxmin_norm(fit32):
    """lde_optimizer.py:80-86 on float64 up-casts of the float32 costs."""
    a = fit32.astype(np.float64)
    if np.max(a) != np.min(a):
        return (a - np.min(a)) / (np.max(a) - np.min(a))
    return np.zeros_like(a)


def histogram5(fitness):
    """np.histogram(fitness, BINS)[0] (lde_optimizer.py:150,185) -- numpy's own implementation is the oracle."""
    return np.histogram(fitness, BINS)[0]


def feature(S: State):
    """lde_optimizer.py:145-157 (`S.pop`/`S.fit` already sorted) -> float64 [NP + 2*BINS]."""
    fitness = maxmin_norm(S.fit)
    return np.concatenate([fitness, histogram5(fitness).astype(np.float64), S.hist_sum / S.hist_cnt])


# ----------------------------------------------------------------------------------- the algorithm
def init_population(cfg: Config, problem_eval, optimum, lb, ub, tape):
    """lde_optimizer.py:132-140 -> (State, feature[NP+10])."""
    NP = cfg.NP
    pop = (lb + tape.init_u() * (ub - lb)).astype(np.float32)
    fit = _costs(problem_eval, optimum, pop, tape.noise(0))
    gbest = np.min(fit)
    pop, fit, _ = order_by_f(pop, fit)
    S = State(pop=pop, fit=fit, hist_sum=np.full(BINS, NP / BINS), hist_cnt=1, gbest=gbest, fes=NP, log_index=1,
              curve=[gbest])
    return S, feature(S)


def update(cfg: Config, S: State, action, problem_eval, optimum, lb, ub, tape, aux=None):
    """lde_optimizer.py:159-198.  action: float32 [2*NP] (sf | cr).  -> (State, feature, reward, done)."""
    NP, D = cfg.NP, cfg.dim
    action = np.asarray(action, dtype=np.float32).reshape(-1)
    assert action.shape[0] == 2 * NP
    pop32, fit32, _ = order_by_f(S.pop, S.fit)                   # :160 (identity on a sorted state)
    fitness = maxmin_norm(fit32)
    pop = pop32.astype(np.float64); fit = fit32
    sf, cr = action[:NP], action[NP:]
    # current-to-pbest mutation matrix :88-105
    p_rate = (cfg.P_MIN - P_INI) * S.fes / cfg.maxFEs + P_INI
    high = int(np.ceil(NP * max(0, p_rate)))
    pidx = np.asarray(tape.pbest_idx(high))
    one_minus = (1 - sf).astype(np.float64)                      # float32 subtraction stored into a float64 matrix
    sf64 = sf.astype(np.float64)
    idx = np.arange(NP)
    mu = np.where((pidx != idx)[:, None], one_minus[:, None] * pop + sf64[:, None] * pop[pidx], pop)
    # donor difference :107-130
    r = np.asarray(tape.donors())
    mu = mu + sf64[:, None] * (pop[r[:, 0]] - pop[r[:, 1]])
    # binomial crossover with j_rand :44-48
    u = np.array(tape.cross_u(), dtype=np.float64, copy=True)
    jr = np.asarray(tape.jrand())
    u[idx, jr] = 0.
    trial = np.where(u <= cr.astype(np.float64)[:, None], mu, pop)
    # parent-midpoint repair :31-38
    lo, hi = trial < lb, trial > ub
    trial = np.where(lo, (pop + lb) / 2., np.where(hi, (pop + ub) / 2., trial))
    trial32 = trial.astype(np.float32)
    tfit = _costs(problem_eval, optimum, trial32, tape.noise(0))
    fes = S.fes + NP
    surv = tfit <= fit                                          # :59
    n_pop = np.where(surv[:, None], trial32, pop32)
    n_fit = np.where(surv, tfit, fit)
    bsf, bsf_next = np.float64(fit.min()), np.float64(n_fit.min())
    with np.errstate(divide='ignore', invalid='ignore'):
        reward = (bsf - bsf_next) / bsf                          # :173
    if optimum is None:
        done = fes >= cfg.maxFEs
    else:
        done = bool(fes >= cfg.maxFEs or n_fit.min() <= 1e-8)
    hist_pre = histogram5(fitness)                              # :183-187 pre-update histogram
    N = State(pop=n_pop, fit=n_fit, hist_sum=S.hist_sum + hist_pre, hist_cnt=S.hist_cnt + 1, gbest=n_fit.min(),
              fes=fes, log_index=S.log_index, curve=list(S.curve), done=done)
    if N.fes >= N.log_index * cfg.log_interval:
        N.log_index += 1
        N.curve.append(N.gbest)
    if done:
        if len(N.curve) >= cfg.n_logpoint + 1:
            N.curve[-1] = N.gbest
        else:
            N.curve.append(N.gbest)
    if aux is not None:
        aux.update(trial=trial32, trial_fit=tfit, surv=surv, pidx=pidx, donors=r, hist_pre=hist_pre)
    N.pop, N.fit, order = order_by_f(N.pop, N.fit)              # __get_feature :146
    if aux is not None:
        aux['order'] = order
    return N, feature(N), reward, done
