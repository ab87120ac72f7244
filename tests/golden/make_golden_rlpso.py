#!/usr/bin/env python
"""Golden steps of the UNMODIFIED reference `RL_PSO_Optimizer` (src/optimizer/rl_pso_optimizer.py), build container only:
python tests/golden/make_golden_rlpso.py -> tests/golden/rl_pso.npz.  Teacher forcing as in make_golden.py: F32Problem proxy,
state rounded to float32-representable values after every step; 230 consecutive single-particle steps per case (more than
two passes over the population)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refenv  # noqa: E402

_refenv.activate()
import numpy as np  # noqa: E402

import make_golden as mg  # noqa: E402

CASES = [('f1_d10', 1, 10), ('f15_d10', 15, 10), ('f21_d30', 21, 30), ('f110_d10', 110, 10), ('prot_1R0R_1', '1R0R_1', 12)]
N_STEPS = 230
P = '_RL_PSO_Optimizer__'


def rounded(opt):
    pt = getattr(opt, P + 'particles')
    for k in ('current_position', 'velocity', 'pbest_position', 'c_cost', 'pbest', 'gbest_position'):
        pt[k] = mg.f32r(pt[k])
    pt['gbest_val'] = np.float64(np.float32(pt['gbest_val']))
    setattr(opt, P + 'max_cost', np.float64(np.float32(getattr(opt, P + 'max_cost'))))
    return pt


def main():
    from optimizer import RL_PSO_Optimizer
    out = dict(cases=np.array([c[0] for c in CASES]), n_steps=np.int64(N_STEPS))
    ga = np.random.RandomState(31)
    for key, fid, dim in CASES:
        if isinstance(fid, str):
            inner = mg.protein_instance(fid); cfg = mg.ref_config(12, 1000, 5)
        else:
            inner = mg.bbob_instance(fid, dim); cfg = mg.ref_config(dim)
        prob = _refenv.F32Problem(inner)
        opt = RL_PSO_Optimizer(cfg)
        prob.reset()
        seed0 = 9000 + len(key)
        np.random.seed(seed0)
        state = opt.init_population(prob)
        pt = rounded(opt)
        out[f'{key}/fid'] = np.array(str(fid)); out[f'{key}/dim'] = np.int64(dim)
        out[f'{key}/maxFEs'] = np.int64(cfg.maxFEs); out[f'{key}/n_logpoint'] = np.int64(cfg.n_logpoint); out[f'{key}/seed0'] = np.int64(seed0)
        for k, name in (('x', 'current_position'), ('v', 'velocity'), ('pbest_x', 'pbest_position'), ('c_cost', 'c_cost'), ('pbest', 'pbest'),
                        ('gbest_x', 'gbest_position')):
            out[f'{key}/s0/{k}'] = np.asarray(pt[name], dtype=np.float32)
        out[f'{key}/s0/gbest_val'] = np.float64(pt['gbest_val']); out[f'{key}/s0/gbest_idx'] = np.int64(pt['gbest_index'])
        out[f'{key}/s0/max_cost'] = np.float64(getattr(opt, P + 'max_cost')); out[f'{key}/s0/state'] = np.asarray(state, dtype=np.float64)
        rec = {k: [] for k in ('action', 'seed', 'row_x', 'row_v', 'row_pbest_x', 'row_c_cost', 'row_pbest', 'gbest_x', 'gbest_val',
                               'gbest_idx', 'w', 'reward', 'done', 'state', 'fes', 'log_index', 'curve_len', 'curve_last')}
        for t in range(1, N_STEPS + 1):
            action = np.array([ga.rand()], dtype=np.float32)
            if t % 37 == 0:
                action[0] = np.float32(1.2)                               # outside (0, 1): the backbone does not clamp
            seed = seed0 + t
            np.random.seed(seed)
            j = getattr(opt, P + 'cur_index')
            state, reward, done = opt.update(action, prob)
            pt = rounded(opt)
            rec['action'].append(action[0]); rec['seed'].append(seed)
            rec['row_x'].append(pt['current_position'][j].copy()); rec['row_v'].append(pt['velocity'][j].copy())
            rec['row_pbest_x'].append(pt['pbest_position'][j].copy()); rec['row_c_cost'].append(pt['c_cost'][j]); rec['row_pbest'].append(pt['pbest'][j])
            rec['gbest_x'].append(np.array(pt['gbest_position'])); rec['gbest_val'].append(pt['gbest_val']); rec['gbest_idx'].append(pt['gbest_index'])
            rec['w'].append(getattr(opt, P + 'w')); rec['reward'].append(reward); rec['done'].append(done); rec['state'].append(np.asarray(state, dtype=np.float64))
            rec['fes'].append(opt.fes); rec['log_index'].append(opt.log_index); rec['curve_len'].append(len(opt.cost)); rec['curve_last'].append(opt.cost[-1])
        for k, v in rec.items():
            a = np.array(v)
            out[f'{key}/{k}'] = a.astype(np.float32) if k in ('row_x', 'row_v', 'row_pbest_x', 'row_c_cost', 'row_pbest', 'gbest_x', 'action') else a
    np.savez_compressed(os.path.join(HERE, 'rl_pso.npz'), **out)
    print('rl_pso done', os.path.getsize(os.path.join(HERE, 'rl_pso.npz')))


if __name__ == '__main__':
    main()
