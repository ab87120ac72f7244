#!/usr/bin/env python
"""Golden generations of the UNMODIFIED reference `RL_HPSDE_Optimizer` (src/optimizer/rl_hpsde_optimizer.py), build container only:
python tests/golden/make_golden_hpsde.py -> tests/golden/rl_hpsde.npz.  F32Problem proxy with in-place rounding (every point the reference
evaluates -- trials, walk samples, the initial population -- is float32-representable and stored as such); np.random re-seeded before
every generation so that the oracle can replay the reference's own draws (incl. scipy's cauchy / levy samplers)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refenv  # noqa: E402

_refenv.activate()
import numpy as np  # noqa: E402

import make_golden as mg  # noqa: E402

CASES = [('f1_d10', 1, 10, 400), ('f15_d10', 15, 10, 30), ('f21_d30', 21, 30, 8), ('f110_d10', 110, 10, 30), ('prot_1R0R_1', '1R0R_1', 12, 25)]
P = '_RL_HPSDE_Optimizer__'


def main():
    from optimizer import RL_HPSDE_Optimizer
    out = dict(cases=np.array([c[0] for c in CASES]))
    ga = np.random.RandomState(91)
    for key, fid, dim, n_steps in CASES:
        if isinstance(fid, str):
            inner = mg.protein_instance(fid); cfg = mg.ref_config(12, 12000, 50)
        else:
            inner = mg.bbob_instance(fid, dim); cfg = mg.ref_config(dim)
        prob = _refenv.F32Problem(inner, inplace=True)
        opt = RL_HPSDE_Optimizer(cfg)
        prob.reset()
        seed0 = 5000 + len(key)
        np.random.seed(seed0)
        state = opt.init_population(prob)
        pop = getattr(opt, P + 'population')
        assert np.array_equal(pop.group, mg.f32r(pop.group))
        out[f'{key}/fid'] = np.array(str(fid)); out[f'{key}/dim'] = np.int64(dim)
        out[f'{key}/maxFEs'] = np.int64(cfg.maxFEs); out[f'{key}/n_logpoint'] = np.int64(cfg.n_logpoint); out[f'{key}/seed0'] = np.int64(seed0)
        out[f'{key}/s0/group'] = pop.group.astype(np.float32); out[f'{key}/s0/cost'] = pop.cost.astype(np.float32)
        out[f'{key}/s0/state'] = np.int64(state); out[f'{key}/s0/fes'] = np.int64(opt.fes)
        rec = {k: [] for k in ('action', 'seed', 'NP', 'MF', 'MCr', 'k', 'gbest', 'state', 'reward', 'done', 'fes', 'log_index', 'curve_len',
                               'curve_last')}
        t, done = 0, False
        while t < n_steps and not done:
            t += 1
            action = int(ga.randint(0, 4))
            seed = seed0 + t
            np.random.seed(seed)
            state, reward, done = opt.update(action, prob)
            pop = getattr(opt, P + 'population')
            assert np.array_equal(pop.group, mg.f32r(pop.group))
            rec['action'].append(action); rec['seed'].append(seed); rec['NP'].append(pop.NP); rec['MF'].append(pop.MF.copy()); rec['MCr'].append(pop.MCr.copy())
            rec['k'].append(pop.k); rec['gbest'].append(pop.gbest); rec['state'].append(state); rec['reward'].append(reward); rec['done'].append(done)
            rec['fes'].append(opt.fes); rec['log_index'].append(opt.log_index); rec['curve_len'].append(len(opt.cost)); rec['curve_last'].append(opt.cost[-1])
            out[f'{key}/cost/{t}'] = pop.cost.astype(np.float32)
            out[f'{key}/group/{t}'] = pop.group.astype(np.float32)        # every generation: np.argsort's order of tied costs is not reproducible
        out[f'{key}/n_steps'] = np.int64(t)
        for k, v in rec.items():
            out[f'{key}/{k}'] = np.array(v)
        print(key, 'steps', t, 'done', done, 'fes', opt.fes, 'NP', pop.NP, 'states', sorted(set(rec['state'])), 'reward mean', float(np.mean(rec['reward'])))
    np.savez_compressed(os.path.join(HERE, 'rl_hpsde.npz'), **out)
    print('rl_hpsde done', os.path.getsize(os.path.join(HERE, 'rl_hpsde.npz')))


if __name__ == '__main__':
    main()
