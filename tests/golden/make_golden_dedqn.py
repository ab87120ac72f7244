#!/usr/bin/env python
"""Golden steps of the UNMODIFIED reference `DEDQN_Optimizer` (src/optimizer/dedqn_optimizer.py), build container only:
python tests/golden/make_golden_dedqn.py -> tests/golden/dedqn.npz.  Teacher forcing as in make_golden.py: F32Problem proxy, the
population rounded IN PLACE to float32-representable values after every step (in place, so that the reference's `__gbest` view of a
population row stays a view); 130 consecutive single-individual steps per case (more than one pass over the population), np.random
re-seeded before every step so that the oracle can replay the reference's own draws."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refenv  # noqa: E402

_refenv.activate()
import numpy as np  # noqa: E402

import make_golden as mg  # noqa: E402

CASES = [('f1_d10', 1, 10), ('f15_d10', 15, 10), ('f21_d30', 21, 30), ('f110_d10', 110, 10), ('f124_d10', 124, 10), ('prot_1R0R_1', '1R0R_1', 12)]
N_STEPS = 130
P = '_DEDQN_Optimizer__'


def rounded(opt):
    pop = getattr(opt, P + 'population')
    pop[...] = mg.f32r(pop)
    gb = getattr(opt, P + 'gbest')
    gb[...] = mg.f32r(gb)                     # a view of a population row is already rounded; a re-bound trial array is rounded here
    return pop


def main():
    from optimizer import DEDQN_Optimizer
    out = dict(cases=np.array([c[0] for c in CASES]), n_steps=np.int64(N_STEPS))
    ga = np.random.RandomState(77)
    for key, fid, dim in CASES:
        if isinstance(fid, str):
            inner = mg.protein_instance(fid); cfg = mg.ref_config(12, 30000, 50)
        else:
            inner = mg.bbob_instance(fid, dim); cfg = mg.ref_config(dim, None if key == 'f1_d10' else 30000)   # f1_d10 runs past maxFEs
        prob = _refenv.F32Problem(inner)
        opt = DEDQN_Optimizer(cfg)
        prob.reset()
        seed0 = 7000 + len(key)
        np.random.seed(seed0)
        state = opt.init_population(prob)
        pop = rounded(opt)
        out[f'{key}/fid'] = np.array(str(fid)); out[f'{key}/dim'] = np.int64(dim)
        out[f'{key}/maxFEs'] = np.int64(cfg.maxFEs); out[f'{key}/n_logpoint'] = np.int64(cfg.n_logpoint); out[f'{key}/seed0'] = np.int64(seed0)
        out[f'{key}/s0/pop'] = np.asarray(pop, dtype=np.float32); out[f'{key}/s0/cost'] = np.asarray(getattr(opt, P + 'cost'), dtype=np.float32)
        out[f'{key}/s0/gbest'] = np.asarray(getattr(opt, P + 'gbest'), dtype=np.float32)
        out[f'{key}/s0/gbest_cost'] = np.float64(getattr(opt, P + 'gbest_cost')); out[f'{key}/s0/state'] = np.asarray(state, dtype=np.float64)
        out[f'{key}/s0/fes'] = np.int64(opt.fes)
        rec = {k: [] for k in ('action', 'seed', 'row', 'row_cost', 'row_survival', 'gbest', 'gbest_cost', 'state', 'reward', 'done', 'fes',
                               'log_index', 'curve_len', 'curve_last', 'replaced')}
        for t in range(1, N_STEPS + 1):
            action = int(ga.randint(0, 3))
            seed = seed0 + t
            np.random.seed(seed)
            j = getattr(opt, P + 'solution_pointer')
            state, reward, done = opt.update(action, prob)
            pop = rounded(opt)
            sv = getattr(opt, P + 'survival')
            rec['action'].append(action); rec['seed'].append(seed); rec['row'].append(pop[j].copy()); rec['row_cost'].append(getattr(opt, P + 'cost')[j])
            rec['row_survival'].append(sv[j]); rec['replaced'].append(sv[j] == 1)
            rec['gbest'].append(np.array(getattr(opt, P + 'gbest'))); rec['gbest_cost'].append(getattr(opt, P + 'gbest_cost'))
            rec['state'].append(np.asarray(state, dtype=np.float64)); rec['reward'].append(reward); rec['done'].append(done)
            rec['fes'].append(opt.fes); rec['log_index'].append(opt.log_index); rec['curve_len'].append(len(opt.cost)); rec['curve_last'].append(opt.cost[-1])
        for k, v in rec.items():
            a = np.array(v)
            out[f'{key}/{k}'] = a.astype(np.float32) if k in ('row', 'row_cost', 'gbest') else a
        out[f'{key}/final_pop'] = np.asarray(pop, dtype=np.float32); out[f'{key}/final_cost'] = np.asarray(getattr(opt, P + 'cost'), dtype=np.float32)
        out[f'{key}/final_survival'] = np.asarray(getattr(opt, P + 'survival'), dtype=np.float64)
        print(key, 'fes', opt.fes, 'replaced', int(np.sum(rec['replaced'])), 'of', N_STEPS, 'state[-1]', rec['state'][-1], 'reward', rec['reward'][-1])
    np.savez_compressed(os.path.join(HERE, 'dedqn.npz'), **out)
    print('dedqn done', os.path.getsize(os.path.join(HERE, 'dedqn.npz')))


if __name__ == '__main__':
    main()
