"""GPU: the batched Tester / rollout drivers write the reference's on-disk layout (tester.py:123-127, 282-352)."""
import argparse
import os
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def _weights(group):
    G = np.load(os.path.join(GOLDEN, 'agents.npz'))
    return {k[len(group) + 1:]: G[k] for k in G.files if k.startswith(group + '/')}


def _config(tmp_path, **kw):
    c = dict(problem='bbob', difficulty='easy', dim=10, maxFEs=2000, n_logpoint=50, log_interval=40, upperbound=5,
             test_log_dir=str(tmp_path / 'test') + '/', rollout_log_dir=str(tmp_path / 'rollout') + '/',
             agent_for_cp=['RLEPSO_Agent', 'LDE_Agent'], t_optimizer_for_cp=['Random_search'])
    c.update(kw)
    return argparse.Namespace(**c)


def test_tester_writes_reference_layout(tmp_path):
    from metabox_b200 import agent as ma, tester as mt
    cfg = _config(tmp_path)
    pol = {'RLEPSO_Agent': ma.load_reference_weights(ma.RLEPSO_Actor(), _weights('rlepso')).cuda().eval(),
           'LDE_Agent': ma.load_reference_weights(ma.LDE_PolicyNet(), _weights('lde')).cuda().eval()}
    t = mt.Tester(cfg, pol)
    out = t.test(n_runs=3, seed=1)
    with open(cfg.test_log_dir + 'test.pkl', 'rb') as f:
        R = pickle.load(f)
    with open(cfg.test_log_dir + 'random_search_baseline.pkl', 'rb') as f:
        B = pickle.load(f)
    assert set(R) == {'cost', 'fes', 'T0', 'T1', 'T2'} and R['T0'] > 0
    names = [str(p) for p in t.test_set.data]
    assert sorted(R['cost']) == sorted(names) and len(names) == 6
    for p in names:
        assert set(R['cost'][p]) == {'RLEPSO_Agent', 'LDE_Agent', 'Random_search'}
        for a in R['cost'][p]:
            runs = R['cost'][p][a]
            assert len(runs) == 3 and all(len(c) == 51 and isinstance(c[0], float) for c in runs)
            assert all(np.all(np.diff(c) <= 0) for c in runs)                       # gbest curves never go up
            assert all(isinstance(f, int) and (f >= 2000 or runs[k][-1] <= 1e-8) for k, f in enumerate(R['fes'][p][a]))
    assert set(R['T1']) == set(R['T2']) == {'RLEPSO_Agent', 'LDE_Agent', 'Random_search'}
    assert len(B['cost']) == 24 and all(list(v) == ['Random_search'] for v in B['cost'].values())
    # learned optimizers beat random search on the sphere-like problems after 2000 FEs (sanity of the whole pipeline)
    p = names[0]
    assert np.mean([c[-1] for c in R['cost'][p]['RLEPSO_Agent']]) < np.mean([c[-1] for c in R['cost'][p]['Random_search']])
    assert out['test'] is R or out['test'] == R


def test_tester_runs_gleet_rl_pso_and_qlpso_backbones_with_callable_policies(tmp_path):
    import torch
    from metabox_b200 import agent as ma, tester as mt
    cfg = _config(tmp_path, agent_for_cp=['GLEET_Agent', 'RL_PSO_Agent', 'QLPSO_Agent'], t_optimizer_for_cp=[], maxFEs=600, log_interval=12)
    pol = {'GLEET_Agent': lambda s: torch.sigmoid(s[..., 1] - s[..., 2]),                 # [B,NP,27] -> [B,NP]
           'RL_PSO_Agent': lambda s: torch.sigmoid(s[:, 0]),                              # [B,2D] -> [B]
           'QLPSO_Agent': ma.QLPSO_QTable(np.array([[1., 0, 0, 2], [0, 1, 0, 0], [0, 0, 1, 0], [2, 0, 0, 1]])).cuda()}   # [B] int -> [B] int
    R = mt.Tester(cfg, pol).test(n_runs=2, random_search_baseline=False)['test']
    for p in R['cost']:
        assert set(R['cost'][p]) == {'GLEET_Agent', 'RL_PSO_Agent', 'QLPSO_Agent'}
        for a in R['cost'][p]:
            assert len(R['cost'][p][a]) == 2 and all(len(c) == 51 for c in R['cost'][p][a])
            assert all(f >= 600 or c[-1] <= 1e-8 for f, c in zip(R['fes'][p][a], R['cost'][p][a]))


def test_rollout_writes_reference_layout(tmp_path):
    from metabox_b200 import agent as ma, tester as mt
    cfg = _config(tmp_path)
    nets = [ma.RLEPSO_Actor().cuda().eval(), ma.load_reference_weights(ma.RLEPSO_Actor(), _weights('rlepso')).cuda().eval()]
    R = mt.rollout(cfg, {'RLEPSO_Agent': nets}, n_runs=2)
    with open(cfg.rollout_log_dir + 'rollout.pkl', 'rb') as f:
        P = pickle.load(f)
    assert set(P) == {'cost', 'fes', 'return'} and len(P['cost']) == 18          # bbob easy train set
    for p in P['cost']:
        ck = P['cost'][p]['RLEPSO_Agent']
        assert len(ck) == 2 and all(len(runs) == 2 and len(runs[0]) == 51 for runs in ck)
        assert all(isinstance(r, float) for runs in P['return'][p]['RLEPSO_Agent'] for r in runs)
    with pytest.raises(ValueError):
        mt.rollout(cfg, {'RL_DAS_Agent': nets})                 # no B200 backbone of that name
    with pytest.raises(ValueError):
        mt.rollout(cfg, {'NRLPSO_Agent': nets})                 # a backbone that exists, handed the policy of another one
    assert R.keys() == P.keys()
