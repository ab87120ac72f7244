"""Oracle: random-number providers.  TEST INFRASTRUCTURE ONLY.

Two interchangeable providers feed the optimizer oracles:

* `NumpyDraws`   -- forwards to the process-global `np.random` with exactly the calls (function,
  argument form, shape, order) the reference makes, so that after `np.random.seed(s)` an oracle
  step consumes the MT19937 stream identically to the reference step (`src/tester.py:198`).
* `PhiloxDraws`  -- Philox4x32-10 keyed by (seed) with counter (index, generation, lane, tag); this is
  the numpy twin of `metabox_b200/csrc/mbx_philox.cuh`, so production (non-tape) CUDA runs can be
  replayed bit-exactly on the CPU.

Both expose the same *tagged* API; `NumpyDraws` ignores the tags, `PhiloxDraws` ignores call order.
"""
from __future__ import annotations

import numpy as np

# ---- draw-group tags (must match mbx_philox.cuh) -------------------------------------------------
TAG_INIT = 1        # per (i,d): w0 -> position u, w1 -> velocity u
TAG_PD = 2          # per (i,d): w0 -> clpso u, w1,w2 -> tournament picks, w3 -> fdr u     (RLEPSO)
TAG_P = 3           # per i    : w0 -> rand1, w1 -> rand2, w2 -> re-init test u               (RLEPSO)
TAG_REINIT = 4      # per (i,d): w0 -> position u, w1 -> velocity u                           (RLEPSO)
TAG_NOISE = 5       # per i    : noise draws of evaluation pass 0 (main); +1 for the re-init pass
TAG_NOISE2 = 6
TAG_LDE_P = 7       # per i    : w0 -> p-best pick, w1,w2 -> r1,r2 proposals, w3 -> j_rand        (LDE)
TAG_LDE_PD = 8      # per (i,d): w0 -> crossover u                                              (LDE)
TAG_LDE_REJ = 9     # per (i,k): rejection redraws for donor indices                            (LDE)
TAG_DDQN = 10       # per step : w0..w3 + second block -> 5 indices r, j_rand                   (DE-DDQN)
TAG_DDQN_D = 11     # per d    : crossover u                                                    (DE-DDQN)
TAG_GLEET_P = 12    # per i    : w0 -> rand1, w1 -> rand2                                         (GLEET)
TAG_RLPSO = 13      # per step : w0 -> rand1                                                      (RL_PSO)
TAG_QLPSO = 14      # per step : w0, w1 -> the two scalars of the velocity; init: per i w0 & 3 -> q-state (QLPSO)
GEN_INIT = 0xFFFFFFFF

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    """Vectorised Philox4x32-`rounds`.  All inputs broadcastable uint32; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & _MASK for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(rounds):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return [c.astype(np.uint32) for c in (c0, c1, c2, c3)]


def u01(w):
    """uint32 -> float64 uniform in [0,1): w * 2^-32."""
    return w.astype(np.float64) * (1.0 / 4294967296.0)


def bounded(w, high):
    """uint32 -> integer in [0, high): (w * high) >> 32."""
    return ((w.astype(np.uint64) * np.uint64(high)) >> np.uint64(32)).astype(np.int64)


def box_muller(wa, wb):
    """Two uint32 words -> two float64 standard normals (r cos t, r sin t), u_a in (0,1]."""
    ua = (wa.astype(np.float64) + 1.0) * (1.0 / 4294967296.0)
    ub = wb.astype(np.float64) * (1.0 / 4294967296.0)
    r = np.sqrt(-2.0 * np.log(ua))
    t = 6.283185307179586 * ub
    return r * np.cos(t), r * np.sin(t)


class PhiloxDraws:
    """Counter-based provider for one lane (population) at one generation."""

    def __init__(self, seed, lane, gen):
        self.k0 = seed & 0xFFFFFFFF
        self.k1 = (seed >> 32) & 0xFFFFFFFF
        self.lane = lane
        self.gen = gen

    def words(self, tag, n):
        """4 words for each of n consecutive indices -> uint32 [n,4]."""
        idx = np.arange(n, dtype=np.uint64)
        w = philox4x32(idx, np.uint64(self.gen), np.uint64(self.lane), np.uint64(tag), self.k0, self.k1)
        return np.stack(w, axis=-1)

    def words_at(self, tag, idx):
        w = philox4x32(np.asarray(idx, dtype=np.uint64), np.uint64(self.gen), np.uint64(self.lane),
                       np.uint64(tag), self.k0, self.k1)
        return np.stack(w, axis=-1)


class NumpyDraws:
    """Marker provider: the optimizer oracles call np.random directly, in reference order."""
    pass
