"""GPU: GLEET fused step parity -- CUDA (through the C ABI) vs golden reference trajectories and the oracle (tape mode and
Philox mode), plus invariants at the benchmark size."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
ATOL27 = np.where(np.arange(27) % 9 == 8, 5e-6, 1e-9)        # cosine feature: see tests/test_oracle_golden.py


def _close(a, b, atol, tag):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    bad = np.abs(a - b) > atol + 2e-6 * np.abs(b)
    assert not bad.any(), f'{tag}: {int(bad.sum())} mismatches, worst {np.abs(a - b)[bad].max()}'


def _check(st, feat, S, state, tag):
    for k in ('x', 'v', 'pbest_x', 'gbest_x'):
        np.testing.assert_array_equal(st[k], getattr(S, k), err_msg=f'{tag} {k}')
    for k in ('c_cost', 'pbest'):
        np.testing.assert_allclose(st[k], getattr(S, k), rtol=1e-6, atol=1e-10, err_msg=f'{tag} {k}')
    np.testing.assert_array_equal(st['pni'], S.pni, err_msg=f'{tag} pni')
    assert int(st['gbest_idx']) == S.gbest_idx and int(st['no_improve']) == S.no_improve, tag
    assert int(st['fes']) == S.fes and int(st['log_index']) == S.log_index and float(st['w']) == S.w, tag
    _close(st['pbest_feature'], S.pbest_feature, ATOL27[None, :9], f'{tag} pbest_feature')
    _close(st['gbest_feature'], S.gbest_feature, ATOL27[:9], f'{tag} gbest_feature')
    _close(feat, state, ATOL27[None, :], f'{tag} state')
    np.testing.assert_allclose(st['curve'][:int(st['curve_len'])], np.array(S.curve, dtype=np.float64), rtol=1e-6, atol=1e-10)
    assert bool(st['done']) == bool(S.done), tag


def _problem(fid, dim):
    from test_gpu_rlepso import _problem as p
    return p(fid, dim)


def test_golden_trajectories_tape_mode():
    """Each golden step of the real GLEET_Optimizer: the GPU starts from the reference's (float32) state, replays the
    reference's draws from a tape, and must land on the reference's next state."""
    import torch
    from metabox_b200 import optimizer as mo, problem as mp
    from oracle import gleet as og
    G = np.load(os.path.join(GOLDEN, 'gleet.npz'))
    for key in G['cases']:
        fid, dim = str(G[f'{key}/fid']), int(G[f'{key}/dim'])
        P, ev, optimum, lb, ub, nc, I = _problem(fid, dim)
        pack = mp.InstancePack([P])
        cfg = og.Config(dim=dim, maxFEs=int(G[f'{key}/maxFEs']), n_logpoint=int(G[f'{key}/n_logpoint']))
        eng = mo.GLEET_Engine(pack, [0], cfg.maxFEs, cfg.n_logpoint)

        def tape_for(seed, init):
            np.random.seed(seed)
            t = {}
            if init:
                t['init_pos_u'] = np.random.random_sample((1, cfg.NP, dim)); t['init_vel_u'] = np.random.random_sample((1, cfg.NP, dim))
            else:
                t['rand1'] = np.random.rand(cfg.NP, 1)[:, 0][None]; t['rand2'] = np.random.rand(cfg.NP, 1)[:, 0][None]
            nz = np.zeros((1, 3, cfg.NP))
            nu, nn = nc
            for j in range(nu):
                nz[0, j] = np.random.rand(cfg.NP)
            for j in range(nn):
                nz[0, nu + j] = np.random.randn(cfg.NP)
            t['noise'] = nz
            return t

        feat = eng.init_population(tape_for(int(G[f'{key}/seed0']), True))
        from test_oracle_golden import _gleet_state_from
        S0 = _gleet_state_from(G, f'{key}/s0/')
        S0.done = False
        _check(eng.read_state(0), feat[0].cpu().numpy(), S0, G[f'{key}/s0/state'], f'{key} init')
        for t in range(1, int(G[f'{key}/n_steps']) + 1):
            prev = _gleet_state_from(G, f'{key}/s{t - 1}/')
            eng.load_state(0, x=prev.x, v=prev.v, pbest_x=prev.pbest_x, c_cost=prev.c_cost, pbest=prev.pbest, pni=prev.pni,
                           gbest_x=prev.gbest_x, gbest_val=prev.gbest_val, gbest_idx=prev.gbest_idx, no_improve=prev.no_improve,
                           w=prev.w, max_cost=prev.max_cost, pbest_feature=prev.pbest_feature, gbest_feature=prev.gbest_feature,
                           fes=prev.fes, log_index=prev.log_index)
            eng.curve[0, :len(prev.curve)] = torch.as_tensor(np.array(prev.curve, dtype=np.float32)).to(eng.device)
            eng.curve_len[0] = len(prev.curve)
            a = torch.as_tensor(G[f'{key}/s{t}/action'][None]).to(eng.device)
            feat, reward, done = eng.update(a, tape_for(int(G[f'{key}/s{t}/seed']), False))
            N = _gleet_state_from(G, f'{key}/s{t}/')
            N.done = bool(G[f'{key}/s{t}/done'])
            _check(eng.read_state(0), feat[0].cpu().numpy(), N, G[f'{key}/s{t}/state'], f'{key} step {t}')
            np.testing.assert_allclose(float(reward[0]), float(G[f'{key}/s{t}/reward']), rtol=2e-6, atol=1e-7)


def test_philox_production_path_vs_oracle():
    """Production RNG: each CUDA generation (Philox in-kernel) is replayed by the oracle with the numpy Philox twin from the
    GPU's previous state -- positions bit-exact over 30 generations, 6 lanes."""
    import torch
    from metabox_b200 import optimizer as mo, problem as mp
    from oracle import bbob as ob, gleet as og
    dim, maxFEs = 10, 20000
    fids = [1, 3, 9, 17, 21, 24]
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', dim, 5)
    probs = {p.fid: p for p in tr.data + te.data}
    otr, ote = ob.make_suite('bbob', dim, 5.0)
    oin = {i.fid: i for i in otr + ote}
    pack = mp.InstancePack([probs[f] for f in fids])
    seed = 0xABCDEF77
    eng = mo.GLEET_Engine(pack, list(range(len(fids))), maxFEs, 50, seed=seed, lane_base=3)
    cfg = og.Config(dim=dim, maxFEs=maxFEs)
    feat = eng.init_population().cpu().numpy()
    for b, f in enumerate(fids):
        I = oin[f]
        ev = lambda x, draws, I=I: ob.func_true(I, x)
        S, state = og.init_population(cfg, ev, I.optimum, -5., 5., og.PhiloxTape(cfg.NP, dim, seed, 3 + b, 0xFFFFFFFF))
        _check(eng.read_state(b), feat[b], S, state, f'init lane {b}')
    g = np.random.RandomState(3)
    for gen in range(30):
        action = g.rand(len(fids), cfg.NP).astype(np.float32)
        prev = [eng.read_state(b) for b in range(len(fids))]
        feat, reward, done = eng.update(torch.as_tensor(action).to(eng.device))
        feat, reward = feat.cpu().numpy(), reward.cpu().numpy()
        for b, f in enumerate(fids):
            I = oin[f]
            ev = lambda x, draws, I=I: ob.func_true(I, x)
            p = prev[b]
            S = og.State(x=p['x'], v=p['v'], pbest_x=p['pbest_x'], c_cost=p['c_cost'], pbest=p['pbest'], pni=p['pni'],
                         gbest_x=p['gbest_x'], gbest_val=np.float32(p['gbest_val']), gbest_idx=int(p['gbest_idx']),
                         no_improve=int(p['no_improve']), w=float(p['w']), max_cost=np.float32(p['max_cost']),
                         pbest_feature=p['pbest_feature'], gbest_feature=p['gbest_feature'], fes=int(p['fes']),
                         log_index=int(p['log_index']), curve=list(p['curve'][:int(p['curve_len'])]))
            N, state, r, d = og.update(cfg, S, action[b], ev, I.optimum, -5., 5., og.PhiloxTape(cfg.NP, dim, seed, 3 + b, gen))
            _check(eng.read_state(b), feat[b], N, state, f'gen {gen} lane {b} (fid {f})')
            np.testing.assert_allclose(reward[b], r, rtol=2e-6, atol=1e-7)


def test_dropin_and_invariants_at_benchmark_size():
    import argparse
    import torch
    from metabox_b200 import optimizer as mo, problem as mp, environment as me
    cfg = argparse.Namespace(dim=10, maxFEs=3000, n_logpoint=50, log_interval=60)
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', 10, 5)
    env = me.PBO_Env(te.data[0], mo.GLEET_Optimizer(cfg))
    np.random.seed(0)
    state = env.reset()
    assert state.shape == (100, 27) and state.dtype == np.float64
    done, n = False, 0
    while not done:
        state, reward, done = env.step(np.random.rand(100).astype(np.float32))
        assert reward >= 0 and state.shape == (100, 27)
        n += 1
    assert env.optimizer.fes >= 3000 or env.optimizer.cost[-1] <= 1e-8
    # BBOB dim 30, 24 functions x 8 runs
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', 30, 5)
    probs = tr.data + te.data
    pack = mp.InstancePack(probs)
    eng = mo.GLEET_Engine(pack, np.repeat(np.arange(24), 8), 60000, 50, seed=5)
    eng.init_population()
    prev_pbest, prev_gbest = eng.pbest.clone(), eng.gbest_val.clone()
    g = torch.Generator(device='cpu').manual_seed(0)
    for gen in range(20):
        feat, reward, done = eng.update(torch.rand(eng.B, 100, generator=g).to(eng.device))
        assert torch.all(eng.x.abs() <= 5.0) and torch.all(eng.v.abs() <= 1.0) and torch.all(eng.pbest <= prev_pbest)
        assert torch.equal(eng.gbest_val, torch.minimum(prev_gbest, eng.pbest.min(dim=1).values))
        assert torch.all(torch.isfinite(feat)) and torch.all(reward >= 0)
        assert torch.all(feat[:, :, 8].abs() <= 1.0 + 1e-6) and torch.all(feat[:, :, 6] >= 0)
        prev_pbest, prev_gbest = eng.pbest.clone(), eng.gbest_val.clone()
    with pytest.raises(ValueError):
        eng.update(torch.zeros(eng.B, 35))
