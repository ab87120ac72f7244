#!/usr/bin/env python
"""Golden values of the UNMODIFIED reference's landscape-feature functions (dedqn_optimizer.py cal_fdc / cal_rie / cal_acf /
cal_nop; rl_hpsde_optimizer.py __DFDC / __DRIE), build container only:  python tests/golden/make_golden_landscape.py"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refenv  # noqa: E402

_refenv.activate()
import numpy as np  # noqa: E402

from optimizer import dedqn_optimizer as rd  # noqa: E402
from optimizer.rl_hpsde_optimizer import RL_HPSDE_Optimizer as RH  # noqa: E402


def hpsde(sample, cost, ref):
    """call the private methods on a stand-in object; the reference returns booleans (or raises outside [-1, 1] / [0, 1])"""
    me = types.SimpleNamespace()
    setattr(me, '_RL_HPSDE_Optimizer__population', types.SimpleNamespace(gbest_solution=ref))
    setattr(me, '_RL_HPSDE_Optimizer__rw_steps', len(cost) - 1)
    try:
        a = int(RH._RL_HPSDE_Optimizer__DFDC(me, sample, cost))
    except ValueError:
        a = -1
    try:
        b = int(RH._RL_HPSDE_Optimizer__DRIE(me, cost))
    except ValueError:
        b = -1
    return a, b


def main():
    g = np.random.RandomState(2024)
    out, keys = {}, []
    for S, D, kind in [(100, 10, 'walk'), (100, 30, 'walk'), (101, 30, 'sphere'), (50, 12, 'rugged'), (100, 30, 'steps'),
                       (64, 10, 'walk'), (100, 10, 'sphere'), (200, 30, 'rugged'), (100, 30, 'plateau')]:
        for rep in range(3):
            key = f'{kind}_S{S}_D{D}_{rep}'
            if kind in ('walk', 'rugged', 'steps', 'plateau'):
                # dedqn_optimizer.py:79-89 random walk in the unit cube, scaled to [-5, 5]
                p = g.rand(D); pts = [p]
                for _ in range(S - 1):
                    p = (p + g.rand(D) * (0.1 if kind != 'walk' else 1.0)) % 1
                    pts.append(p)
                x = -5 + 10 * np.array(pts)
            else:
                x = g.uniform(-5, 5, (S, D))
            if kind == 'rugged':
                c = np.sum(x ** 2, -1) + 10 * np.sum(np.cos(7 * x), -1)
            elif kind == 'steps':
                c = np.floor(np.sum(np.abs(x), -1) / 4)                      # many equal neighbours
            elif kind == 'plateau':
                c = np.maximum(np.sum(x ** 2, -1), 260.)                     # flat stretches
            else:
                c = np.sum((x - 1.0) ** 2, -1)
            ref = x[g.randint(S)] * 0.5
            a, b = hpsde(x, c, ref)
            out[key + '/x'], out[key + '/c'], out[key + '/ref'] = x, c, ref
            out[key + '/dedqn'] = np.array([rd.cal_fdc(x, c), rd.cal_rie(c), rd.cal_acf(c), rd.cal_nop(x, c)])
            out[key + '/hpsde_bool'] = np.array([a, b])
            keys.append(key)
    out['cases'] = np.array(keys)
    np.savez_compressed(os.path.join(HERE, 'landscape.npz'), **out)
    print('landscape done', len(keys))


if __name__ == '__main__':
    main()
