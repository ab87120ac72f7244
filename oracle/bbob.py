"""Oracle: BBOB (F1-F24) and noisy BBOB (F101-F130) instances and fitness.  TEST INFRASTRUCTURE ONLY.

numpy float64 restatement of the reference `src/problem/bbob.py`.  Structure is table driven
(one `Instance` record + one function per *base shape*) instead of the reference's class tree, but
every arithmetic chain keeps the reference's operation order so that elementwise IEEE results are
bit-identical; sums / libm calls agree to ~1e-15 relative.  Pinned against the reference by
`tests/golden/bbob_*.npz` (see `tests/golden/make_golden.py`).

Reference anchors (file:line under /root/reference/src/problem/):
  sr_func bbob.py:6-8, rotate_gen :11-28, ctor/optimum :31-45, osc :51-67, asy :70-82, pen :85-93,
  noise mixins :96-146, the function classes :149-890, dataset factory :904-952.
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np

NOISE_NONE, NOISE_GAUSS, NOISE_UNIFORM, NOISE_CAUCHY = 0, 1, 2, 3

# fid -> (base shape id, display name).  Names follow the reference `__str__` of each class.
_NOISELESS_NAMES = {
    1: 'Sphere', 2: 'Ellipsoidal', 3: 'Rastrigin', 4: 'Buche_Rastrigin', 5: 'Linear_Slope',
    6: 'Attractive_Sector', 7: 'Step_Ellipsoidal', 8: 'Rosenbrock_original', 9: 'Rosenbrock_rotated',
    10: 'Ellipsoidal_high_cond', 11: 'Discus', 12: 'Bent_Cigar', 13: 'Sharp_Ridge', 14: 'Different_Powers',
    15: 'Rastrigin_F15', 16: 'Weierstrass', 17: 'Schaffers', 18: 'Schaffers_high_cond',
    19: 'Composite_Grie_rosen', 20: 'Schwefel', 21: 'Gallagher_101Peaks', 22: 'Gallagher_21Peaks',
    23: 'Katsuura', 24: 'Lunacek_bi_Rastrigin',
}

# noisy triples: first fid of the triple -> (base shape, stem, moderate?)   bbob.py:170-207,336-353,...
_NOISY_TRIPLES = {
    101: (1, 'Sphere_moderate', True), 104: (8, 'Rosenbrock_moderate', True),
    107: (1, 'Sphere', False), 110: (8, 'Rosenbrock', False), 113: (7, 'Step_Ellipsoidal', False),
    116: (10, 'Ellipsoidal', False), 119: (14, 'Different_Powers', False), 122: (17, 'Schaffers', False),
    125: (19, 'Composite_Grie_rosen', False), 128: (21, 'Gallagher_101Peaks', False),
}
_NOISE_SUFFIX = {NOISE_GAUSS: 'gauss', NOISE_UNIFORM: 'uniform', NOISE_CAUCHY: 'cauchy'}

BBOB_IDS = list(range(1, 25))
NOISY_IDS = list(range(101, 131))
SMALL_SET = {'bbob': [1, 5, 6, 10, 15, 20],
             'bbob-noisy': [101, 105, 115, 116, 117, 119, 120, 125]}   # bbob.py:918,921


@dataclasses.dataclass
class Instance:
    fid: int
    base: int
    name: str
    dim: int
    lb: float
    ub: float
    bias: float
    shift: np.ndarray                 # ctor-final shift == optimum location   [D]
    R: np.ndarray                     # ctor-final rotation                    [D,D]
    M2: Optional[np.ndarray] = None   # Q_rotate (F7) / linearTF (F9,F15-F19)  [D,D]
    y: Optional[np.ndarray] = None    # Gallagher peaks   [P,D]
    C: Optional[np.ndarray] = None    # Gallagher scales  [P,D]
    w: Optional[np.ndarray] = None    # Gallagher weights [P]
    bh_coef: float = 0.0              # boundaryHandling(x) = bh_coef * pen(x)
    cond: float = 0.0                 # _Ellipsoidal.condition
    factor: float = 0.0               # _Composite_Grie_rosen.factor
    noise: int = NOISE_NONE
    n_alpha: float = 0.0              # uniform_alpha / cauchy_alpha
    n_beta: float = 0.0               # gauss_beta / uniform_beta
    n_p: float = 0.0                  # cauchy_p
    optimum: float = 0.0
    FES: int = 0

    def __str__(self):
        return self.name


# ----------------------------------------------------------------------------- shared transforms
def rotation(dim, rng=np.random):
    """Random rotation by successive Householder reflections (bbob.py:11-28); consumes
    `rng.normal(size=dim-n+1)` for n = 1..dim-1 in that order."""
    H = np.eye(dim)
    signs = np.ones((dim,))
    for n in range(1, dim):
        v = rng.normal(size=(dim - n + 1,))
        signs[n - 1] = np.sign(v[0])
        v[0] -= signs[n - 1] * np.sqrt((v * v).sum())
        block = np.eye(dim - n + 1) - 2. * np.outer(v, v) / (v * v).sum()
        step = np.eye(dim)
        step[n - 1:, n - 1:] = block
        H = np.dot(H, step)
    signs[-1] = (-1) ** (1 - (dim % 2)) * signs.prod()
    return (signs * H.T).T


def shift_rotate(x, s, R):
    """z = (R @ (x - s)^T)^T  (bbob.py:6-8)."""
    d = x[:, :s.shape[-1]] - s
    return np.matmul(R, d.transpose()).transpose()


def t_osz(x):
    """Oscillation transform (bbob.py:51-67).  For x>0 the reference uses sin(t)+sin(0.79 t), for
    x<0 sin(0.55 t)+sin(0.31 t) with t = log|x|/0.1; 0 stays 0."""
    out = np.array(x, dtype=np.float64, copy=True)
    pos = x > 0.
    t = np.log(x[pos]) / 0.1
    out[pos] = np.exp(t + 0.49 * (np.sin(t) + np.sin(0.79 * t))) ** 0.1
    neg = x < 0.
    t = np.log(-x[neg]) / 0.1
    out[neg] = -np.exp(t + 0.49 * (np.sin(0.55 * t) + np.sin(0.31 * t))) ** 0.1
    return out


def t_asy(x, beta):
    """Asymmetry transform (bbob.py:70-82): x_i^(1 + beta * i/(D-1) * sqrt(x_i)) for x_i > 0."""
    n, dim = x.shape
    frac = np.broadcast_to(np.linspace(0, 1, dim).reshape(1, -1), (n, dim))
    out = x.copy()
    pos = x > 0.
    out[pos] = out[pos] ** (1. + beta * frac[pos] * np.sqrt(out[pos]))
    return out


def penalty(x, ub):
    """sum_i max(0, |x_i| - ub)^2  (bbob.py:85-93)."""
    return np.sum(np.maximum(0., np.abs(x) - ub) ** 2, axis=-1)


def _rosen_sum(z):
    return np.sum(100 * (z[:, :-1] ** 2 - z[:, 1:]) ** 2 + (z[:, :-1] - 1) ** 2, axis=-1)


def _rastrigin_sum(z, dim):
    return 10. * (dim - np.sum(np.cos(2. * np.pi * z), axis=-1)) + np.sum(z ** 2, axis=-1)


# ----------------------------------------------------------------------------- base shapes
def _boundary(I, x):
    # NoisyProblem.boundaryHandling bbob.py:104-105 and the per-class overrides
    if I.bh_coef == 0.0:
        return 0.
    if I.bh_coef == 1.0:
        return penalty(x, I.ub)
    return I.bh_coef * penalty(x, I.ub)


def _sphere(I, x):                     # bbob.py:156-159
    z = shift_rotate(x, I.shift, I.R)
    return np.sum(z ** 2, axis=-1) + I.bias + _boundary(I, x)


def _ellipsoidal_f2(I, x):             # bbob.py:220-226
    z = t_osz(shift_rotate(x, I.shift, I.R))
    i = np.arange(I.dim)
    return np.sum(np.power(10, 6 * i / (I.dim - 1)) * (z ** 2), -1) + I.bias


def _rastrigin_f3(I, x):               # bbob.py:234,240-243
    scales = (10. ** 0.5) ** np.linspace(0, 1, I.dim)
    z = scales * t_asy(t_osz(shift_rotate(x, I.shift, I.R)), beta=0.2)
    return _rastrigin_sum(z, I.dim) + I.bias


def _buche_f4(I, x):                   # bbob.py:252,258-265
    scales = (10. ** 0.5) ** np.linspace(0, 1, I.dim)
    z = t_osz(shift_rotate(x, I.shift, I.R))
    ev = z[:, ::2]
    ev[ev > 0.] *= 10.                 # view write-through, as in the reference
    z *= scales
    return 10 * (I.dim - np.sum(np.cos(2 * np.pi * z), axis=-1)) + np.sum(z ** 2, axis=-1) \
        + 100 * penalty(x, I.ub) + I.bias


def _linear_slope_f5(I, x):            # bbob.py:281-287
    z = x.copy()
    over = (x * I.shift) > (I.ub ** 2)
    z[over] = np.sign(z[over]) * I.ub
    s = np.sign(I.shift) * (10 ** np.linspace(0, 1, I.dim))
    return np.sum(I.ub * np.abs(s) - z * s, axis=-1) + I.bias


def _attractive_f6(I, x):              # bbob.py:302-307
    z = shift_rotate(x, I.shift, I.R)
    z = np.ascontiguousarray(z)
    hit = (z * I.shift) > 0.
    z[hit] *= 100.
    return t_osz(np.sum(z ** 2, -1)) ** 0.9 + I.bias


def _step_ellipsoidal(I, x):           # bbob.py:320-325
    zh = shift_rotate(x, I.shift, I.R)
    zt = np.where(np.abs(zh) > 0.5, np.floor(0.5 + zh), np.floor(0.5 + 10. * zh) / 10.)
    z = np.matmul(zt, I.M2.T)
    return 0.1 * np.maximum(np.abs(zh[:, 0]) / 1e4,
                            np.sum(100 ** np.linspace(0, 1, I.dim) * (z ** 2), axis=-1)) \
        + _boundary(I, x) + I.bias


def _rosenbrock(I, x):                 # bbob.py:365-368
    z = max(1., I.dim ** 0.5 / 8.) * shift_rotate(x, I.shift, I.R) + 1
    return _rosen_sum(z) + I.bias + _boundary(I, x)


def _rosenbrock_rot_f9(I, x):          # bbob.py:432-435
    z = np.matmul(x, I.M2.T) + 0.5
    return _rosen_sum(z) + I.bias


def _ellipsoidal(I, x):                # bbob.py:447-453
    z = t_osz(shift_rotate(x, I.shift, I.R))
    i = np.arange(I.dim)
    return np.sum((I.cond ** (i / (I.dim - 1))) * (z ** 2), -1) + I.bias + _boundary(I, x)


def _discus_f11(I, x):                 # bbob.py:498-502
    z = t_osz(shift_rotate(x, I.shift, I.R))
    return np.power(10, 6) * (z[:, 0] ** 2) + np.sum(z[:, 1:] ** 2, -1) + I.bias


def _bent_cigar_f12(I, x):             # bbob.py:517-522
    z = t_asy(shift_rotate(x, I.shift, I.R), beta=0.5)
    z = np.matmul(z, I.R.T)
    return z[:, 0] ** 2 + np.sum(np.power(10, 6) * (z[:, 1:] ** 2), -1) + I.bias


def _sharp_ridge_f13(I, x):            # bbob.py:537-540
    z = shift_rotate(x, I.shift, I.R)
    return z[:, 0] ** 2. + 100. * np.sqrt(np.sum(z[:, 1:] ** 2., axis=-1)) + I.bias


def _dif_powers(I, x):                 # bbob.py:550-554
    z = shift_rotate(x, I.shift, I.R)
    i = np.arange(I.dim)
    return np.power(np.sum(np.power(np.fabs(z), 2 + 4 * i / max(1, I.dim - 1)), -1), 0.5) \
        + I.bias + _boundary(I, x)


def _rastrigin_f15(I, x):              # bbob.py:597-602
    z = t_asy(t_osz(shift_rotate(x, I.shift, I.R)), beta=0.2)
    z = np.matmul(z, I.M2.T)
    return _rastrigin_sum(z, I.dim) + I.bias


def _weierstrass_f16(I, x):            # bbob.py:612-614,620-625
    aK = 0.5 ** np.arange(12)
    bK = 3.0 ** np.arange(12)
    f0 = np.sum(aK * np.cos(np.pi * bK))
    z = np.matmul(t_osz(shift_rotate(x, I.shift, I.R)), I.M2.T)
    inner = np.sum(aK * np.cos(np.matmul(2 * np.pi * (z[:, :, None] + 0.5), bK[None, :])), axis=-1)
    return 10 * np.power(np.mean(inner, axis=-1) - f0, 3) + 10 / I.dim * penalty(x, I.ub) + I.bias


def _schaffers(I, x):                  # bbob.py:639-645
    z = np.matmul(t_asy(shift_rotate(x, I.shift, I.R), beta=0.5), I.M2.T)
    s = np.sqrt(z[:, :-1] ** 2 + z[:, 1:] ** 2)
    return np.power(1 / (I.dim - 1) * np.sum(np.sqrt(s) * (np.power(np.sin(50 * np.power(s, 0.2)), 2) + 1),
                                             axis=-1), 2) + _boundary(I, x) + I.bias


def _grie_rosen(I, x):                 # bbob.py:701-705
    z = np.matmul(x, I.M2.T) + 0.5
    s = 100. * (z[:, :-1] ** 2 - z[:, 1:]) ** 2 + (1. - z[:, :-1]) ** 2
    return I.factor + I.factor * np.sum(s / 4000. - np.cos(s), axis=-1) / (I.dim - 1.) + I.bias \
        + _boundary(I, x)


def _schwefel_f20(I, x):               # bbob.py:751-759
    tmp = 2 * np.abs(I.shift)
    scales = (10 ** 0.5) ** np.linspace(0, 1, I.dim)
    z = 2 * np.sign(I.shift) * x
    z[:, 1:] += 0.25 * (z[:, :-1] - tmp[:-1])       # right-hand side uses the OLD z (not recursive)
    z = 100. * (scales * (z - tmp) + tmp)
    return 4.189828872724339 - 0.01 * np.mean(z * np.sin(np.sqrt(np.abs(z))), axis=-1) \
        + 100 * penalty(z / 100, I.ub) + I.bias


def _gallagher(I, x):                  # bbob.py:795-799
    P = I.y.shape[0]
    z = np.matmul(np.expand_dims(x, axis=1).repeat(P, axis=1) - I.y, I.R.T)       # [NP,P,D]
    g = np.max(I.w * np.exp((-0.5 / I.dim) * np.sum(I.C * (z ** 2), axis=-1)), axis=-1)
    return t_osz(10 - g) ** 2 + I.bias + _boundary(I, x)


def _katsuura_f23(I, x):               # bbob.py:855-866
    z = shift_rotate(x, I.shift, I.R)
    pw = np.power(I.dim, 1.2)
    two_j = np.repeat(np.power(np.ones((1, 32)) * 2, np.arange(1, 33)), x.shape[0], 0)
    prod = np.ones(x.shape[0])
    for i in range(I.dim):
        t = two_j * np.repeat(z[:, i, None], 32, 1)
        acc = np.sum(np.fabs(t - np.floor(t + 0.5)) / two_j, -1)
        prod *= np.power(1 + (i + 1) * acc, 10 / pw)
    c = 10 / I.dim / I.dim
    return prod * c - c + penalty(x, I.ub) + I.bias


def _lunacek_f24(I, x):                # bbob.py:874,883-890
    mu0 = 2.5 / 5 * I.ub
    xh = 2. * np.sign(I.shift) * x
    z = np.matmul(xh - mu0, I.R.T)
    s = 1. - 1. / (2. * np.sqrt(I.dim + 20.) - 8.2)
    mu1 = -np.sqrt((mu0 ** 2 - 1) / s)
    return np.minimum(np.sum((xh - mu0) ** 2., axis=-1), I.dim + s * np.sum((xh - mu1) ** 2., axis=-1)) \
        + 10. * (I.dim - np.sum(np.cos(2. * np.pi * z), axis=-1)) + 1e4 * penalty(x, I.ub) + I.bias


_BASE = {1: _sphere, 2: _ellipsoidal_f2, 3: _rastrigin_f3, 4: _buche_f4, 5: _linear_slope_f5,
         6: _attractive_f6, 7: _step_ellipsoidal, 8: _rosenbrock, 9: _rosenbrock_rot_f9, 10: _ellipsoidal,
         11: _discus_f11, 12: _bent_cigar_f12, 13: _sharp_ridge_f13, 14: _dif_powers, 15: _rastrigin_f15,
         16: _weierstrass_f16, 17: _schaffers, 18: _schaffers, 19: _grie_rosen, 20: _schwefel_f20,
         21: _gallagher, 22: _gallagher, 23: _katsuura_f23, 24: _lunacek_f24}


def func_true(I: Instance, x):
    """Noise-free f(x) including bias and boundary handling; x [NP,D] float64 -> [NP]."""
    x = np.asarray(x, dtype=np.float64)
    I.FES += x.shape[0]
    return _BASE[I.base](I, x)


def apply_noise(I: Instance, ftrue, rng=np.random, draws=None):
    """Noise models bbob.py:108-146.  `draws` (optional) is a tuple of pre-drawn arrays in the
    reference's draw order -- Gauss: (randn,), Uniform: (rand, rand), Cauchy: (rand, randn, randn) --
    so a recorded tape can be replayed; otherwise they are taken from `rng` in that order."""
    if I.noise == NOISE_NONE:
        return ftrue
    ftrue = np.asarray(ftrue, dtype=np.float64)
    shp = ftrue.shape
    u = ftrue - I.optimum
    if I.noise == NOISE_GAUSS:
        n0 = rng.randn(*shp) if draws is None else draws[0]
        v = u * np.exp(I.n_beta * n0)
    elif I.noise == NOISE_UNIFORM:
        if draws is None:
            u1 = rng.rand(*shp)
            u2 = rng.rand(*shp)
        else:
            u1, u2 = draws
        v = u * (u1 ** I.n_beta) * \
            np.maximum(1., (1e9 / (u + 1e-99)) ** (I.n_alpha * (0.49 + 1. / I.dim) * u2))
    else:
        if draws is None:
            u1 = rng.rand(*shp)
            n1 = rng.randn(*shp)
            n2 = rng.randn(*shp)
        else:
            u1, n1, n2 = draws
        v = u + I.n_alpha * np.maximum(0., 1e3 + (u1 < I.n_p) * n1 / (np.abs(n2) + 1e-199))
    return np.where(u >= 1e-8, v + I.optimum + 1.01 * 1e-8, ftrue)


def noise_draw_count(I: Instance):
    """(#uniform arrays, #normal arrays) one eval consumes, in order (uniforms first for Cauchy)."""
    return {NOISE_NONE: (0, 0), NOISE_GAUSS: (0, 1), NOISE_UNIFORM: (2, 0), NOISE_CAUCHY: (1, 2)}[I.noise]


def evaluate(I: Instance, x, rng=np.random, draws=None):
    """`Basic_Problem.eval` (basic_problem.py:12-34) incl. the noisy override (bbob.py:100-102):
    1-D x -> scalar, 2-D -> [NP], N-D -> flattened to [-1, D]."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        f = func_true(I, x.reshape(1, -1))[0]
    elif x.ndim == 2:
        f = func_true(I, x)
    else:
        f = func_true(I, x.reshape(-1, x.shape[-1]))
    return apply_noise(I, f, rng, draws)


# ----------------------------------------------------------------------------- instance factory
def make_instance(fid, dim, shift, rotate, bias, lb, ub, rng=np.random) -> Instance:
    """Per-class constructors (bbob.py:149-890): applies the ctor transforms to (shift, rotate) and
    draws the extra constructor randomness from `rng` in the reference's order."""
    noise, n_alpha, n_beta, n_p = NOISE_NONE, 0.0, 0.0, 0.0
    bh = 0.0
    if fid >= 101:
        first = 101 + 3 * ((fid - 101) // 3)
        base, stem, moderate = _NOISY_TRIPLES[first]
        noise = (NOISE_GAUSS, NOISE_UNIFORM, NOISE_CAUCHY)[fid - first]
        name = stem + '_' + _NOISE_SUFFIX[noise]
        if noise == NOISE_GAUSS:
            n_beta = 0.01 if moderate else 1.
        elif noise == NOISE_UNIFORM:
            n_alpha = n_beta = 0.01 if moderate else 1.
        else:
            n_alpha, n_p = (0.01, 0.05) if moderate else (1., 0.2)
        bh = 100.0                                   # NoisyProblem.boundaryHandling bbob.py:104-105
    else:
        base, name = fid, _NOISELESS_NAMES[fid]
        bh = {7: 1.0, 17: 10.0, 18: 10.0, 21: 1.0, 22: 1.0}.get(fid, 0.0)

    I = Instance(fid=fid, base=base, name=name, dim=dim, lb=lb, ub=ub, bias=bias,
                 shift=shift, R=rotate, bh_coef=bh, noise=noise, n_alpha=n_alpha, n_beta=n_beta, n_p=n_p)
    lam10 = (10. ** 0.5) ** np.linspace(0, 1, dim)

    if base == 4:                                   # bbob.py:251
        shift[::2] = np.abs(shift[::2])
    elif base == 5:                                 # bbob.py:273-275
        s = np.sign(shift)
        s[s == 0.] = rng.choice([-1., 1.], size=(s == 0.).sum())
        I.shift = s * ub
    elif base in (6, 13):                           # bbob.py:295-296, 530-531
        I.R = np.matmul(np.matmul(rotation(dim, rng), np.diag(lam10)), rotate)
    elif base == 7:                                 # bbob.py:315-317
        I.R = np.matmul(np.diag(lam10), rotate)
        I.M2 = rotation(dim, rng)
    elif base == 8:                                 # bbob.py:361-362
        shift *= 0.75
        I.R = np.eye(dim)
    elif base == 9:                                 # bbob.py:424-426
        scale = max(1., dim ** 0.5 / 8.)
        I.M2 = scale * rotate
        I.shift = np.matmul(0.5 * np.ones(dim), I.M2) / (scale ** 2)
    elif base == 10:                                # bbob.py:457,466
        I.cond = 1e6 if fid == 10 else 1e4
    elif base == 15:                                # bbob.py:590-591
        I.M2 = np.matmul(np.matmul(rotate, np.diag(lam10)), rotation(dim, rng))
    elif base == 16:                                # bbob.py:610-611
        sc = (0.01 ** 0.5) ** np.linspace(0, 1, dim)
        I.M2 = np.matmul(np.matmul(rotate, np.diag(sc)), rotation(dim, rng))
    elif base in (17, 18):                          # bbob.py:635-636
        cond = 10. if base == 17 else 1000.
        sc = (cond ** 0.5) ** np.linspace(0, 1, dim)
        I.M2 = np.matmul(np.diag(sc), rotation(dim, rng))
    elif base == 19:                                # bbob.py:696-698
        scale = max(1., dim ** 0.5 / 8.)
        I.M2 = scale * rotate
        I.shift = np.matmul(0.5 * np.ones(dim) / (scale ** 2.), I.M2)
        I.factor = 10. if fid == 19 else 1.
    elif base == 20:                                # bbob.py:745
        I.shift = 0.5 * 4.2096874633 * rng.choice([-1., 1.], size=dim)
    elif base in (21, 22):                          # bbob.py:768-791
        P = 101 if base == 21 else 21
        shrink, g_alpha = (1., 1e3) if P == 101 else (0.98, 1e6)
        I.y = shrink * (rng.rand(P, dim) * (ub - lb) + lb)
        I.y[0] = shift * shrink
        I.shift = I.y[0]
        sa = 1000 ** rng.permutation(np.linspace(0, 1, P - 1))
        sa = np.insert(sa, obj=0, values=np.sqrt(g_alpha))
        I.C = np.vstack([rng.permutation(sa[i] ** np.linspace(-0.5, 0.5, dim)) for i in range(P)])
        I.w = np.insert(np.linspace(1.1, 9.1, P - 1), 0, 10.)
    elif base == 23:                                # bbob.py:848-849
        sc = (100. ** 0.5) ** np.linspace(0, 1, dim)
        I.R = np.matmul(np.matmul(rotation(dim, rng), np.diag(sc)), rotate)
    elif base == 24:                                # bbob.py:874-877
        mu0 = 2.5 / 5 * ub
        I.shift = rng.choice([-1., 1.], size=dim) * mu0 / 2
        sc = (100 ** 0.5) ** np.linspace(0, 1, dim)
        I.R = np.matmul(np.matmul(rotation(dim, rng), np.diag(sc)), rotate)

    I.optimum = func_true(I, I.shift.reshape(1, -1))[0]           # bbob.py:40-42 (noise free)
    I.FES = 0
    return I


def make_suite(suit, dim, upperbound=5., difficulty='easy', instance_seed=3849, rng=np.random):
    """`BBOB_Dataset.get_datasets` (bbob.py:904-952): returns (train_list, test_list) of Instances.
    Draw order per function id: shift = 0.8*U(lb,ub)^D, rotation, bias = randint(1,26)*100, ctor."""
    if suit == 'bbob':
        ids = BBOB_IDS
    elif suit == 'bbob-noisy':
        ids = NOISY_IDS
    else:
        raise ValueError(f'{suit} function suit is invalid or is not supported yet.')
    if difficulty not in ('easy', 'difficult'):
        raise ValueError(f'{difficulty} difficulty is invalid.')
    small = SMALL_SET[suit]
    if instance_seed > 0:
        rng.seed(instance_seed)
    assert upperbound >= 5.
    ub, lb = upperbound, -upperbound
    train, test = [], []
    for fid in ids:
        shift = 0.8 * (rng.random(dim) * (ub - lb) + lb)
        H = rotation(dim, rng)
        bias = rng.randint(1, 26) * 100
        inst = make_instance(fid, dim, shift, H, bias, lb, ub, rng)
        in_train = (difficulty == 'easy' and fid not in small) or (difficulty == 'difficult' and fid in small)
        (train if in_train else test).append(inst)
    return train, test
