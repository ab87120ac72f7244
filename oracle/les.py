"""Oracle: LES backbone (a learned evolution strategy: self-attention over the 16 individuals' fitness features gives recombination
weights, a per-dimension MLP gives the learning rates of mean and step size).  TEST INFRASTRUCTURE ONLY.

numpy / torch-CPU restatement of `src/optimizer/les_optimizer.py:7-183`, one GENERATION per call (the reference runs all generations
inside one `update`; its `skip_step` option makes single generations observable).  float64 state like the reference (the population is
16 x dim; nothing is rounded to float32 here), float32 inside the two small networks like the reference (`torch.FloatTensor`).

Reference behaviour reproduced on purpose: costs are RAW function values (the optimum is never subtracted, :71, :139); the "rank" feature
is `argsort(costs) / NP - 0.5` (the sorting permutation, not the rank, :89); `improved = costs < gbest` is evaluated after gbest already
contains this generation (:90, :141); the evolution paths use `(1 - a) P + a (drift - P)` (:97-100); the new mean is
`(1 - a) mu + a sum((x - mu) w)` without adding mu back (:128-129).

Pinned by `tests/golden/les.npz` (the real `LES_Optimizer` stepped with skip_step = 1, tests/golden/make_golden_les.py).
"""
from __future__ import annotations

import dataclasses

import numpy as np
import torch

NP_ = 16                                                       # les_optimizer.py:55
SIGMA_RATIO = 0.2                                              # :56
ALPHAS = (0.1, 0.5, 0.9)                                       # :52
TIMESTAMP = np.array([1, 3, 10, 30, 50, 100, 250, 500, 750, 1000, 1250, 1500, 2000])       # :53
N_ATTN, N_MLP = 68, 178                                        # parameters of SelfAttn (:21-24) / LrNet (:36-38)


@dataclasses.dataclass
class State:
    parents: np.ndarray        # [NP,D] f64
    cost: np.ndarray           # [NP]   f64 raw function values
    t: int
    gbest: float
    Pc: np.ndarray             # [3,D]
    Ps: np.ndarray             # [3,D]
    mu: np.ndarray             # [D]
    sigma: np.ndarray          # [D]
    fes: int
    log_index: int
    curve: list
    done: bool = False

    def copy(self):
        kw = {f.name: (getattr(self, f.name).copy() if isinstance(getattr(self, f.name), np.ndarray) else getattr(self, f.name))
              for f in dataclasses.fields(self)}
        kw['curve'] = list(self.curve)
        return State(**kw)


def _split(vec, shapes):
    out, p = [], 0
    for shp in shapes:
        n = int(np.prod(shp))
        out.append(torch.as_tensor(np.asarray(vec[p:p + n], dtype=np.float32)).reshape(shp))
        p += n
    assert p == len(vec)
    return out


@torch.no_grad()
def attention_weights(attn_vec, cost, gbest):
    """:86-91 + SelfAttn (:19-32): cost [NP] f64 -> recombination weights W [NP] (float32 values)."""
    Wq, bq, Wk, bk, Wv, bv = _split(attn_vec, [(8, 3), (8,), (8, 3), (8,), (1, 3), (1,)])
    z = (cost - np.mean(cost)) / (np.std(cost) + 1e-8)
    rank = np.argsort(cost) / NP_ - 0.5
    X = torch.from_numpy(np.vstack([z, rank, cost < gbest]).T).to(torch.float32)
    lin = torch.nn.functional.linear                                    # (nn.Linear's own kernel: the float32 rounding must match)
    Q, K, V = lin(X, Wq, bq), lin(X, Wk, bk), lin(X, Wv, bv)
    score = torch.softmax(torch.matmul(Q, K.T) / np.sqrt(8), dim=-1)
    return torch.softmax(torch.matmul(score, V), dim=0).squeeze().numpy()


@torch.no_grad()
def learning_rates(mlp_vec, S, W, dim):
    """:93-110 + LrNet (:34-45) -> (alpha [D,2] float32 values, new Pc [3,D], new Ps [3,D])."""
    W1, b1, W2, b2 = _split(mlp_vec, [(8, 19), (8,), (2, 8), (2,)])
    drift = np.sum((S.parents - S.mu) * W[:, None], axis=0)
    drift_s = np.sum((S.parents - S.mu) / S.sigma * W[:, None], axis=0)
    Pc = np.vstack([(1 - a) * S.Pc[i] + a * (drift - S.Pc[i]) for i, a in enumerate(ALPHAS)])
    Ps = np.vstack([(1 - a) * S.Ps[i] + a * (drift_s - S.Ps[i]) for i, a in enumerate(ALPHAS)])
    rho = np.tanh(S.t / TIMESTAMP - 1)[None, :].repeat(dim, axis=0)
    F = torch.from_numpy(np.hstack([Pc.T, Ps.T, rho])).to(torch.float32)
    lin = torch.nn.functional.linear
    return torch.sigmoid(lin(lin(F, W1, b1), W2, b2)).numpy(), Pc, Ps


def init_population(dim, maxFEs, ev, lb, ub, n_logpoint=50):
    """:59-82 (draws from the global np.random in the reference's order)."""
    mu = lb + (ub - lb) * np.random.rand(dim)
    sigma = np.ones(dim) * ub * SIGMA_RATIO
    pop = np.clip(np.random.normal(mu, sigma, (NP_, dim)), lb, ub)
    cost = np.asarray(ev(pop, None), dtype=np.float64)
    return State(parents=pop, cost=cost, t=0, gbest=float(np.min(cost)), Pc=np.zeros((3, dim)), Ps=np.zeros((3, dim)), mu=mu, sigma=sigma,
                 fes=NP_, log_index=1, curve=[float(np.min(cost))])


def generation(S: State, attn_vec, mlp_vec, dim, maxFEs, log_interval, ev, optimum, lb, ub, normal=None):
    """One pass of the loop body (:118-160) -> new State (its `done` = the reference's `is_end` without skip_step).  `normal`: the
    [NP, D] standard-normal draws to use instead of np.random.normal (replay of the batched engine's tape)."""
    S = S.copy()
    W = attention_weights(attn_vec, S.cost, S.gbest)
    alpha, Pc, Ps = learning_rates(mlp_vec, S, W, dim)
    a_mu, a_sigma = alpha[:, 0], alpha[:, 1]
    mu = (1 - a_mu) * S.mu + a_mu * np.sum((S.parents - S.mu) * W[:, None], axis=0)
    sigma = (1 - a_sigma) * S.sigma + a_sigma * np.sqrt(np.sum((S.parents - S.mu) ** 2 * W[:, None], axis=0))
    raw = np.random.normal(mu, sigma, (NP_, dim)) if normal is None else mu + sigma * normal
    pop = np.clip(raw, lb, ub)
    cost = np.asarray(ev(pop, None), dtype=np.float64)
    S.fes += NP_
    S.gbest = float(np.min([np.min(cost), S.gbest]))
    S.parents, S.cost, S.t, S.Pc, S.Ps, S.mu, S.sigma = pop, cost, S.t + 1, Pc, Ps, mu, sigma
    S.done = bool(S.fes >= maxFEs) if optimum is None else bool(S.fes >= maxFEs or S.gbest <= 1e-8)
    if S.fes >= S.log_index * log_interval:
        S.log_index += 1
        S.curve.append(S.gbest)
    return S
