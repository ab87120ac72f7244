#!/usr/bin/env python
"""Golden steps of the UNMODIFIED reference `QLPSO_Optimizer` (src/optimizer/qlpso_optimizer.py), build container only:
python tests/golden/make_golden_qlpso.py -> tests/golden/qlpso.npz.  Teacher forcing as in make_golden.py: F32Problem proxy,
state rounded to float32-representable values after every step (the diversity scalar is recomputed from the rounded
population, as the engine's storage contract defines it); 100 consecutive single-particle steps per case (more than three
passes over the 30 particles), actions cycling through all four neighbourhood sizes."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refenv  # noqa: E402

_refenv.activate()
import numpy as np  # noqa: E402

import make_golden as mg  # noqa: E402

CASES = [('f1_d10', 1, 10), ('f15_d10', 15, 10), ('f21_d30', 21, 30), ('f110_d10', 110, 10), ('f7_d30', 7, 30), ('prot_1R0R_1', '1R0R_1', 12)]
N_STEPS = 100
P = '_QLPSO_Optimizer__'


def rounded(opt):
    for k in ('population', 'velocity', 'pbest', 'cost'):
        setattr(opt, P + k, mg.f32r(getattr(opt, P + k)))
    setattr(opt, P + 'gbest_cost', np.float64(np.float32(getattr(opt, P + 'gbest_cost'))))
    pop = getattr(opt, P + 'population')
    setattr(opt, P + 'diversity', np.mean(np.sqrt(np.sum(np.square(pop - np.mean(pop, 0)), 1))))


def main():
    import warnings
    warnings.simplefilter('ignore', DeprecationWarning)
    from optimizer import QLPSO_Optimizer
    out = dict(cases=np.array([c[0] for c in CASES]), n_steps=np.int64(N_STEPS))
    ga = np.random.RandomState(47)
    for key, fid, dim in CASES:
        if isinstance(fid, str):
            inner = mg.protein_instance(fid); cfg = mg.ref_config(12, 1000, 5)
        else:
            inner = mg.bbob_instance(fid, dim); cfg = mg.ref_config(dim)
        prob = _refenv.F32Problem(inner)
        opt = QLPSO_Optimizer(cfg)
        prob.reset()
        seed0 = 12000 + len(key)
        np.random.seed(seed0)
        qs = opt.init_population(prob)
        rounded(opt)
        g = lambda k: getattr(opt, P + k)
        out[f'{key}/fid'] = np.array(str(fid)); out[f'{key}/dim'] = np.int64(dim)
        out[f'{key}/maxFEs'] = np.int64(cfg.maxFEs); out[f'{key}/n_logpoint'] = np.int64(cfg.n_logpoint); out[f'{key}/seed0'] = np.int64(seed0)
        out[f'{key}/s0/x'] = g('population').astype(np.float32); out[f'{key}/s0/v'] = g('velocity').astype(np.float32)
        out[f'{key}/s0/pbest_x'] = g('pbest').astype(np.float32); out[f'{key}/s0/cost'] = g('cost').astype(np.float32)
        out[f'{key}/s0/gbest_cost'] = np.float64(g('gbest_cost')); out[f'{key}/s0/diversity'] = np.float64(g('diversity'))
        out[f'{key}/s0/q_state'] = np.asarray(g('state'), dtype=np.int32); out[f'{key}/s0/ret_state'] = np.int64(qs)
        rec = {k: [] for k in ('action', 'seed', 'row_x', 'row_v', 'row_pbest_x', 'row_cost', 'gbest_cost', 'diversity', 'reward', 'done',
                               'next_state', 'fes', 'log_index', 'curve_len', 'curve_last')}
        for t in range(1, N_STEPS + 1):
            action = np.array([(t + t // 30) % 4 if t % 7 else ga.randint(0, 4)])      # shape (1,) like np.random.choice(size=1)
            seed = seed0 + t
            np.random.seed(seed)
            j = g('solution_pointer')
            nxt, reward, done = opt.update(action, prob)
            rounded(opt)
            rec['action'].append(int(action[0])); rec['seed'].append(seed)
            rec['row_x'].append(g('population')[j].copy()); rec['row_v'].append(g('velocity')[j].copy()); rec['row_pbest_x'].append(g('pbest')[j].copy())
            rec['row_cost'].append(g('cost')[j]); rec['gbest_cost'].append(g('gbest_cost')); rec['diversity'].append(g('diversity'))
            rec['reward'].append(int(reward)); rec['done'].append(bool(done)); rec['next_state'].append(int(np.squeeze(nxt)))
            rec['fes'].append(opt.fes); rec['log_index'].append(opt.log_index); rec['curve_len'].append(len(opt.cost)); rec['curve_last'].append(float(np.squeeze(opt.cost[-1])))
        for k, v in rec.items():
            a = np.array(v)
            out[f'{key}/{k}'] = a.astype(np.float32) if k in ('row_x', 'row_v', 'row_pbest_x', 'row_cost') else a
    np.savez_compressed(os.path.join(HERE, 'qlpso.npz'), **out)
    print('qlpso done', os.path.getsize(os.path.join(HERE, 'qlpso.npz')))


if __name__ == '__main__':
    main()
