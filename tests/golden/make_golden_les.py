#!/usr/bin/env python
"""Golden generations of the UNMODIFIED reference `LES_Optimizer` (src/optimizer/les_optimizer.py), build container only:
python tests/golden/make_golden_les.py -> tests/golden/les.npz.  The reference runs a whole episode inside one `update`; its `skip_step`
option (= 1 here) returns after every generation, so each one is observable.  The attention / learning-rate network parameters are seeded
random vectors (they ARE the action of this backbone); np.random is re-seeded before every generation."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refenv  # noqa: E402

_refenv.activate()
import numpy as np  # noqa: E402

import make_golden as mg  # noqa: E402

CASES = [('f1_d10', 1, 10), ('f15_d10', 15, 10), ('f21_d30', 21, 30), ('prot_1R0R_1', '1R0R_1', 12)]
N_GEN = 40


def main():
    from optimizer import LES_Optimizer
    g = np.random.RandomState(17)
    attn, mlp = (g.randn(68) * 0.5).astype(np.float32), (g.randn(178) * 0.3).astype(np.float32)
    out = dict(cases=np.array([c[0] for c in CASES]), n_gen=np.int64(N_GEN), attn=attn, mlp=mlp)
    for key, fid, dim in CASES:
        if isinstance(fid, str):
            prob = mg.protein_instance(fid); cfg = mg.ref_config(12, 1000, 5)
        else:
            prob = mg.bbob_instance(fid, dim); cfg = mg.ref_config(dim)
        opt = LES_Optimizer(cfg)
        prob.reset()
        seed0 = 1100 + len(key)
        np.random.seed(seed0)
        opt.init_population(prob)
        E = opt.evolution_info
        out[f'{key}/fid'] = np.array(str(fid)); out[f'{key}/dim'] = np.int64(dim); out[f'{key}/maxFEs'] = np.int64(cfg.maxFEs)
        out[f'{key}/log_interval'] = np.int64(cfg.log_interval); out[f'{key}/seed0'] = np.int64(seed0)
        out[f'{key}/s0/parents'] = E['parents']; out[f'{key}/s0/cost'] = E['parents_cost']; out[f'{key}/s0/mu'] = E['mu']; out[f'{key}/s0/sigma'] = E['sigma']
        rec = {k: [] for k in ('seed', 'parents', 'cost', 'mu', 'sigma', 'Pc', 'Ps', 'gbest', 'fes', 'log_index', 'curve_len', 'ret')}
        for t in range(1, N_GEN + 1):
            seed = seed0 + t
            np.random.seed(seed)
            gb, ret, is_end, _ = opt.update({'attn': attn, 'mlp': mlp, 'skip_step': 1}, prob)
            E = opt.evolution_info
            rec['seed'].append(seed); rec['parents'].append(E['parents'].copy()); rec['cost'].append(np.asarray(E['parents_cost']).copy())
            rec['mu'].append(E['mu'].copy()); rec['sigma'].append(E['sigma'].copy()); rec['Pc'].append(np.asarray(E['Pc']).copy()); rec['Ps'].append(np.asarray(E['Ps']).copy())
            rec['gbest'].append(gb); rec['fes'].append(opt.FEs); rec['log_index'].append(opt.log_index); rec['curve_len'].append(len(opt.cost)); rec['ret'].append(ret)
        for k, v in rec.items():
            out[f'{key}/{k}'] = np.array(v)
        print(key, 'gbest', rec['gbest'][0], '->', rec['gbest'][-1], 'sigma mean', float(np.mean(rec['sigma'][-1])), 'fes', opt.FEs)
    np.savez_compressed(os.path.join(HERE, 'les.npz'), **out)
    print('les done', os.path.getsize(os.path.join(HERE, 'les.npz')))


if __name__ == '__main__':
    main()
