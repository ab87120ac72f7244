"""GPU: Random_search (BASELINE config C1) -- whole episodes in one launch vs the reference's golden episodes (tape replay)
and vs the oracle with the Philox twin (production RNG), incl. pop = 50 and a noisy function."""
import argparse
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def _cfg(dim, **kw):
    return argparse.Namespace(dim=dim, maxFEs=2000 * dim, n_logpoint=50, log_interval=2000 * dim // 50, **kw)


def _prob(fid, dim):
    from metabox_b200 import problem as mp
    suit = 'bbob' if fid < 100 else 'bbob-noisy'
    tr, te = mp.BBOB_Dataset.get_datasets(suit, dim, 5)
    return [p for p in tr.data + te.data if p.fid == fid][0]


@pytest.mark.parametrize('fid,dim', [(1, 10), (15, 10), (8, 30)])
def test_golden_episodes_tape_mode(fid, dim):
    """np.random.seed(5) draws of the reference episode replayed through the tape: fes equal, cost curve within 1e-6."""
    from metabox_b200 import optimizer as mo, problem as mp
    G = np.load(os.path.join(GOLDEN, 'rs.npz'))
    P = _prob(fid, dim)
    pack = mp.InstancePack([P])
    n_iter = int(G[f'f32/{fid}_{dim}/fes']) // 100
    np.random.seed(5)
    tape = np.stack([np.random.random_sample((100, dim)) for _ in range(n_iter)])[None]       # [1, n_iter, NP, D]
    res = mo.Random_search(_cfg(dim)).run_lanes(pack, [0], tape_u=tape)[0]
    assert res['fes'] == int(G[f'f32/{fid}_{dim}/fes'])
    np.testing.assert_allclose(np.array(res['cost']), G[f'f32/{fid}_{dim}/cost'], rtol=1e-6, atol=1e-10)


@pytest.mark.parametrize('fid,dim,NP', [(1, 10, 50), (1, 10, 100), (102, 10, 100), (21, 30, 100)])
def test_philox_production_path_vs_oracle(fid, dim, NP):
    """C1 (Sphere, dim 10, pop 50) and friends with the in-kernel Philox stream: the oracle replays the same counters."""
    from metabox_b200 import optimizer as mo, problem as mp
    from oracle import bbob as ob, random_search as ors
    P = _prob(fid, dim)
    pack = mp.InstancePack([P])
    otr, ote = ob.make_suite('bbob' if fid < 100 else 'bbob-noisy', dim, 5.0)
    I = [i for i in otr + ote if i.fid == fid][0]
    ev = lambda x, draws: ob.apply_noise(I, ob.func_true(I, x), draws=draws)
    cfg = _cfg(dim, rs_NP=NP)
    seed = 77
    res = mo.Random_search(cfg).run_lanes(pack, [0, 0], seed=seed)
    for lane in (0, 1):
        want = ors.run_episode(dim, cfg.maxFEs, 50, cfg.log_interval, ev, I.optimum, -5., 5.,
                               ors.PhiloxTape(NP, dim, seed, lane, noise_kind=I.noise), NP=NP)
        assert res[lane]['fes'] == want['fes'] and len(res[lane]['cost']) == len(want['cost']) == 51
        np.testing.assert_allclose(np.array(res[lane]['cost']), np.array(want['cost'], dtype=np.float64), rtol=1e-6, atol=1e-10)
    assert res[0]['cost'] != res[1]['cost']                       # different lanes, different streams


def test_dropin_run_episode():
    from metabox_b200 import optimizer as mo
    P = _prob(1, 10)
    np.random.seed(0)
    r = mo.Random_search(_cfg(10)).run_episode(P)
    assert r['fes'] == 20000 and len(r['cost']) == 51 and all(a >= b for a, b in zip(r['cost'], r['cost'][1:]))
