"""Oracle: landscape features of a sampled walk.  TEST INFRASTRUCTURE ONLY.

numpy restatement of the feature functions of the DEDQN backbone (`src/optimizer/dedqn_optimizer.py:8-12` cal_fdc,
`:15-51` cal_rie, `:54-61` cal_acf, `:64-76` cal_nop) and of the RL_HPSDE backbone (`src/optimizer/
rl_hpsde_optimizer.py:169-181` DFDC, `:184-216` DRIE; the reference thresholds the two values into booleans, the
oracle returns the raw r / entropy so that the thresholds can be applied by the caller).
Pinned by `tests/golden/landscape.npz` (outputs of the reference's own functions on seeded inputs).
"""
import numpy as np


def cal_fdc(sample, fitness):
    """dedqn_optimizer.py:8-12."""
    best = np.argmin(fitness)
    distance = np.linalg.norm(sample - sample[best], axis=-1)
    cfd = np.mean((fitness - np.mean(fitness)) * (distance - np.mean(distance)))
    return cfd / (np.var(distance) * np.var(fitness) + 1e-6)


def _entropy6(freq):
    return -np.sum(freq * np.log(freq) / np.log(6))


def cal_rie(fitness):
    """dedqn_optimizer.py:15-51 (vectorised; class order: (-1,0) (-1,1) (0,1) (0,-1) (1,-1) else)."""
    f = np.asarray(fitness, dtype=np.float64)
    n = len(f)
    d = f[1:] - f[:-1]
    eps_star = max(0., float(np.max(d))) if n > 1 else 0.
    hs = []
    for k in range(9):
        eps = eps_star / (2 ** k) if k < 8 else 0
        s = np.where(d < -eps, -1, np.where(d > eps, 1, 0))
        a, b = s[:-1], s[1:]
        freq = np.zeros(6)
        freq[0] = np.sum((a == -1) & (b == 0)); freq[1] = np.sum((a == -1) & (b == 1)); freq[2] = np.sum((a == 0) & (b == 1))
        freq[3] = np.sum((a == 0) & (b == -1)); freq[4] = np.sum((a == 1) & (b == -1))
        freq[5] = len(a) - freq[:5].sum()
        freq[freq == 0] = n
        freq /= n
        hs.append(_entropy6(freq))
    return max(hs)


def cal_acf(fitness):
    """dedqn_optimizer.py:54-61."""
    f = np.asarray(fitness, dtype=np.float64)
    avg = np.mean(f)
    a = np.sum((f - avg) ** 2) + 1e-6
    return float(np.sum((f[:-1] - avg) * (f[1:] - avg)) / a)


def cal_nop(sample, fitness):
    """dedqn_optimizer.py:64-76.  np.argsort's order of exactly equal distances is unspecified in the reference (quick
    sort); ties are ordered by index here (stable)."""
    best = np.argmin(fitness)
    distance = np.linalg.norm(sample - sample[best], axis=-1)
    fs = np.asarray(fitness)[np.argsort(distance, kind='stable')]
    return float(np.sum(fs[1:] < fs[:-1])) / len(fitness)


def dfdc_r(sample, cost, ref):
    """rl_hpsde_optimizer.py:169-174: Pearson r between cost and distance to `ref`, first walk point dropped."""
    sample, cost = sample[1:], cost[1:]
    dist = np.linalg.norm(sample - ref, ord=2, axis=-1)
    return float(np.mean((cost - cost.mean()) * (dist - dist.mean())) / (cost.std() * dist.std()))


def drie_r(cost):
    """rl_hpsde_optimizer.py:184-211 with rw_steps = len(cost) - 1."""
    cost = np.asarray(cost, dtype=np.float64)
    rw = len(cost) - 1
    diff = cost[1:] - cost[:rw]
    e_star = np.max(np.abs(diff))
    r = None
    for i, scale in enumerate([0, 1 / 128, 1 / 64, 1 / 32, 1 / 16, 1 / 8, 1 / 4, 1 / 2, 1]):
        symbol = (diff < (-scale * e_star)) * (-1) + ((scale * e_star) < diff) * 1
        a, b = symbol[:rw - 1], symbol[1:rw]
        prob = np.array([np.sum((a == -1) & (b == 0)), np.sum((a == -1) & (b == 1)), np.sum((a == 0) & (b == -1)),
                         np.sum((a == 0) & (b == 1)), np.sum((a == 1) & (b == -1)), np.sum((a == 1) & (b == 0))], dtype=np.float64)
        prob /= rw
        prob[prob < 1e-15] = 1e-15
        e = _entropy6(prob)
        r = e if i == 0 else max(r, e)
    return float(r)


def features(sample, cost, ref=None):
    """-> [fdc, rie, acf, nop, dfdc_r, drie_r] (what mbx_landscape_features writes per lane)."""
    if ref is None:
        ref = sample[np.argmin(cost)]
    return np.array([cal_fdc(sample, cost), cal_rie(cost), cal_acf(cost), cal_nop(sample, cost), dfdc_r(sample, cost, ref), drie_r(cost)])
