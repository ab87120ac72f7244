#!/usr/bin/env python
"""Golden steps of the UNMODIFIED reference `NRLPSO_Optimizer` (src/optimizer/nrlpso_optimizer.py), build container only:
python tests/golden/make_golden_nrlpso.py -> tests/golden/nrlpso.npz.  F32Problem proxy with in-place rounding (every evaluated point --
the initial swarm, the moved particle, the neighbourhood mutants -- is float32-representable and stored as such), velocities rounded in
place after every step; np.random re-seeded before every step; 260 consecutive single-particle steps per case (the pointer wraps twice, so
the neighbourhood snapshots, the stale-table steps and the gbest view aliasing all occur)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refenv  # noqa: E402

_refenv.activate()
import numpy as np  # noqa: E402

import make_golden as mg  # noqa: E402

CASES = [('f1_d10', 1, 10), ('f15_d10', 15, 10), ('f21_d30', 21, 30), ('f110_d10', 110, 10), ('prot_1R0R_1', '1R0R_1', 12)]
N_STEPS = 260
P = '_NRLPSO_Optimizer__'


def main():
    from optimizer import NRLPSO_Optimizer
    out = dict(cases=np.array([c[0] for c in CASES]), n_steps=np.int64(N_STEPS))
    ga = np.random.RandomState(55)
    for key, fid, dim in CASES:
        if isinstance(fid, str):
            inner = mg.protein_instance(fid); cfg = mg.ref_config(12, 1000, 5)
        else:
            inner = mg.bbob_instance(fid, dim); cfg = mg.ref_config(dim)
        prob = _refenv.F32Problem(inner, inplace=True)
        opt = NRLPSO_Optimizer(cfg)
        prob.reset()
        seed0 = 3000 + len(key)
        np.random.seed(seed0)
        state = opt.init_population(prob)
        g = lambda name: getattr(opt, P + name)
        out[f'{key}/fid'] = np.array(str(fid)); out[f'{key}/dim'] = np.int64(dim)
        out[f'{key}/maxFEs'] = np.int64(cfg.maxFEs); out[f'{key}/n_logpoint'] = np.int64(cfg.n_logpoint); out[f'{key}/seed0'] = np.int64(seed0)
        out[f'{key}/s0/pop'] = g('population').astype(np.float32); out[f'{key}/s0/cost'] = g('cost').astype(np.float32)
        out[f'{key}/s0/state'] = np.int64(state); out[f'{key}/s0/r_w'] = np.float64(opt.r_w)
        rec = {k: [] for k in ('action', 'seed', 'row_pop', 'row_vel', 'row_cost', 'row_pbest_pos', 'row_pbest_cost', 'gbest_pos', 'gbest_cost',
                               'stag', 'state', 'reward', 'done', 'fes', 'log_index', 'curve_len', 'curve_last', 'w')}
        for t in range(1, N_STEPS + 1):
            action = int(ga.randint(0, 4))
            seed = seed0 + t
            np.random.seed(seed)
            j = opt.pointer
            state, reward, done = opt.update(action, prob)
            vel = g('velocity'); vel[...] = mg.f32r(vel)
            assert np.array_equal(g('population'), mg.f32r(g('population'))) and np.array_equal(g('pbest_pos'), mg.f32r(g('pbest_pos')))
            rec['action'].append(action); rec['seed'].append(seed); rec['row_pop'].append(g('population')[j].copy()); rec['row_vel'].append(vel[j].copy())
            rec['row_cost'].append(g('cost')[j]); rec['row_pbest_pos'].append(g('pbest_pos')[j].copy()); rec['row_pbest_cost'].append(g('pbest_cost')[j])
            rec['gbest_pos'].append(np.array(g('gbest_pos'))); rec['gbest_cost'].append(g('gbest_cost')); rec['stag'].append(opt.pbest_stag_count[j])
            rec['state'].append(state); rec['reward'].append(reward); rec['done'].append(done); rec['fes'].append(opt.fes)
            rec['log_index'].append(opt.log_index); rec['curve_len'].append(len(opt.cost)); rec['curve_last'].append(opt.cost[-1]); rec['w'].append(opt.w)
            if t % 20 == 0 or t == N_STEPS:
                out[f'{key}/pop/{t}'] = g('population').astype(np.float32); out[f'{key}/cost/{t}'] = g('cost').astype(np.float32)
        for k, v in rec.items():
            a = np.array(v)
            out[f'{key}/{k}'] = a.astype(np.float32) if k in ('row_pop', 'row_vel', 'row_cost', 'row_pbest_pos', 'row_pbest_cost', 'gbest_pos') else a
        print(key, 'fes', opt.fes, 'rewards', {r: int(np.sum(np.array(rec['reward']) == r)) for r in (2, 1, 0, -2)}, 'mutations', (opt.fes - 100 - N_STEPS) // 2)
    np.savez_compressed(os.path.join(HERE, 'nrlpso.npz'), **out)
    print('nrlpso done', os.path.getsize(os.path.join(HERE, 'nrlpso.npz')))


if __name__ == '__main__':
    main()
