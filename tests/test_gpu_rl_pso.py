"""GPU: RL_PSO fused single-particle step -- CUDA (through the C ABI) vs the reference's golden steps (tape replay) and the
oracle with the Philox twin (production RNG)."""
import argparse
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def _problem(fid, dim):
    from test_gpu_rlepso import _problem as p
    return p(fid, dim)


def _noise_tape(nc, n):
    nz = np.zeros((1, 3, n))
    nu, nn = nc
    for j in range(nu):
        nz[0, j] = np.random.rand(n)
    for j in range(nn):
        nz[0, nu + j] = np.random.randn(n)
    return nz


def test_golden_steps_tape_mode():
    """230 consecutive reference steps per case: the GPU carries its OWN state forward (only the float32 costs are
    teacher-forced, they may differ in the last bit) and must reproduce every updated row, gbest and state vector."""
    import torch
    from metabox_b200 import optimizer as mo, problem as mp
    G = np.load(os.path.join(GOLDEN, 'rl_pso.npz'))
    for key in G['cases']:
        fid, dim = str(G[f'{key}/fid']), int(G[f'{key}/dim'])
        P, ev, optimum, lb, ub, nc, I = _problem(fid, dim)
        pack = mp.InstancePack([P])
        eng = mo.RL_PSO_Engine(pack, [0], int(G[f'{key}/maxFEs']), int(G[f'{key}/n_logpoint']))
        np.random.seed(int(G[f'{key}/seed0']))
        tape = dict(init_pos_u=np.random.random_sample((1, 100, dim)), init_vel_u=np.random.random_sample((1, 100, dim)))
        tape['noise'] = _noise_tape(nc, 100)
        st0 = eng.init_population(tape)[0].cpu().numpy()
        s = eng.read_state(0)
        for k in ('x', 'v', 'pbest_x', 'gbest_x'):
            np.testing.assert_array_equal(s[k], G[f'{key}/s0/{k}'], err_msg=f'{key} init {k}')
        np.testing.assert_allclose(s['c_cost'], G[f'{key}/s0/c_cost'], rtol=1e-6, atol=1e-10)
        np.testing.assert_allclose(float(s['max_cost']), float(G[f'{key}/s0/max_cost']), rtol=1e-6)
        np.testing.assert_allclose(st0, G[f'{key}/s0/state'], rtol=1e-6, atol=1e-7)
        eng.load_state(0, c_cost=G[f'{key}/s0/c_cost'], pbest=G[f'{key}/s0/pbest'], gbest_val=np.float32(G[f'{key}/s0/gbest_val']),
                       max_cost=np.float32(G[f'{key}/s0/max_cost']))
        for t in range(1, int(G['n_steps']) + 1):
            k = t - 1
            j = k % 100
            np.random.seed(int(G[f'{key}/seed'][k]))
            tp = dict(rand1=np.array([np.random.rand()]))
            tp['noise'] = _noise_tape(nc, 1)
            a = torch.as_tensor(G[f'{key}/action'][k:k + 1]).to(eng.device)
            state, reward, done = eng.update(a, tp)
            s = eng.read_state(0)
            tag = f'{key} step {t}'
            np.testing.assert_array_equal(s['x'][j], G[f'{key}/row_x'][k], err_msg=tag + ' x')
            np.testing.assert_array_equal(s['v'][j], G[f'{key}/row_v'][k], err_msg=tag + ' v')
            np.testing.assert_array_equal(s['pbest_x'][j], G[f'{key}/row_pbest_x'][k], err_msg=tag + ' pbest_x')
            np.testing.assert_array_equal(s['gbest_x'], G[f'{key}/gbest_x'][k], err_msg=tag + ' gbest_x')
            np.testing.assert_allclose(s['c_cost'][j], G[f'{key}/row_c_cost'][k], rtol=1e-6, atol=1e-10, err_msg=tag)
            assert int(s['gbest_idx']) == int(G[f'{key}/gbest_idx'][k]) and float(s['w']) == float(G[f'{key}/w'][k]), tag
            assert int(s['fes']) == int(G[f'{key}/fes'][k]) and int(s['log_index']) == int(G[f'{key}/log_index'][k]), tag
            assert int(s['curve_len']) == int(G[f'{key}/curve_len'][k]) and bool(s['done']) == bool(G[f'{key}/done'][k]), tag
            np.testing.assert_allclose(state[0].cpu().numpy(), G[f'{key}/state'][k], rtol=1e-6, atol=1e-7, err_msg=tag + ' state')
            np.testing.assert_allclose(float(reward[0]), G[f'{key}/reward'][k], rtol=2e-5, atol=1e-7, err_msg=tag + ' reward')
            eng.c_cost[0, j] = float(G[f'{key}/row_c_cost'][k]); eng.pbest[0, j] = float(G[f'{key}/row_pbest'][k])
            eng.gbest_val[0] = float(np.float32(G[f'{key}/gbest_val'][k]))


def test_philox_production_path_vs_oracle():
    """Production RNG: 260 consecutive CUDA steps (Philox in-kernel) on 5 lanes, each replayed by the oracle with the numpy
    Philox twin from the GPU's previous state: rows bit-exact."""
    import torch
    from metabox_b200 import optimizer as mo, problem as mp
    from oracle import bbob as ob, rl_pso as orp
    dim, maxFEs = 10, 20000
    fids = [1, 8, 15, 21, 24]
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', dim, 5)
    probs = {p.fid: p for p in tr.data + te.data}
    otr, ote = ob.make_suite('bbob', dim, 5.0)
    oin = {i.fid: i for i in otr + ote}
    pack = mp.InstancePack([probs[f] for f in fids])
    seed = 2024
    eng = mo.RL_PSO_Engine(pack, list(range(len(fids))), maxFEs, 50, seed=seed, lane_base=2)
    cfg = orp.Config(dim=dim, maxFEs=maxFEs)
    st = eng.init_population().cpu().numpy()
    for b, f in enumerate(fids):
        I = oin[f]
        ev = lambda x, draws, I=I: ob.func_true(I, x)
        S, state = orp.init_population(cfg, ev, I.optimum, -5., 5., orp.PhiloxTape(100, dim, seed, 2 + b, 0xFFFFFFFF))
        s = eng.read_state(b)
        np.testing.assert_array_equal(s['x'], S.x); np.testing.assert_array_equal(s['gbest_x'], S.gbest_x)
        np.testing.assert_allclose(st[b], state, rtol=1e-6, atol=1e-7)
    g = np.random.RandomState(1)
    for step in range(260):
        action = g.rand(len(fids)).astype(np.float32)
        prev = [eng.read_state(b) for b in range(len(fids))]
        state, reward, done = eng.update(torch.as_tensor(action).to(eng.device))
        state = state.cpu().numpy(); reward = reward.cpu().numpy()
        for b, f in enumerate(fids):
            I = oin[f]
            ev = lambda x, draws, I=I: ob.func_true(I, x)
            p = prev[b]
            S = orp.State(x=p['x'], v=p['v'], pbest_x=p['pbest_x'], c_cost=p['c_cost'], pbest=p['pbest'], gbest_x=p['gbest_x'],
                          gbest_val=np.float32(p['gbest_val']), gbest_idx=int(p['gbest_idx']), w=float(p['w']),
                          max_cost=np.float32(p['max_cost']), cur_index=int(p['cur_index']), fes=int(p['fes']),
                          log_index=int(p['log_index']), curve=list(p['curve'][:int(p['curve_len'])]))
            N, want, r, d = orp.update(cfg, S, action[b], ev, I.optimum, -5., 5., orp.PhiloxTape(100, dim, seed, 2 + b, int(p['fes'])))
            s = eng.read_state(b)
            tag = f'step {step} lane {b}'
            for k in ('x', 'v', 'pbest_x', 'gbest_x'):
                np.testing.assert_array_equal(s[k], getattr(N, k), err_msg=f'{tag} {k}')
            np.testing.assert_allclose(s['c_cost'], N.c_cost, rtol=1e-6, atol=1e-10, err_msg=tag)
            assert int(s['cur_index']) == N.cur_index and int(s['fes']) == N.fes and float(s['w']) == N.w, tag
            np.testing.assert_allclose(state[b], want, rtol=1e-6, atol=1e-7, err_msg=tag)
            np.testing.assert_allclose(reward[b], r, rtol=2e-5, atol=1e-7, err_msg=tag)


def test_dropin_and_batch():
    import torch
    from metabox_b200 import optimizer as mo, problem as mp, environment as me
    cfg = argparse.Namespace(dim=10, maxFEs=600, n_logpoint=50, log_interval=12)
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', 10, 5)
    env = me.PBO_Env(te.data[0], mo.RL_PSO_Optimizer(cfg))
    np.random.seed(0)
    state = env.reset()
    assert state.shape == (20,) and state.dtype == np.float64
    done, n = False, 0
    while not done:
        state, reward, done = env.step(np.array([np.random.rand()], dtype=np.float32))
        n += 1
    assert n == 500 and env.optimizer.fes == 600 and len(env.optimizer.cost) == 51
    # 24 functions x 64 runs, dim 30: 150 lockstep steps
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', 30, 5)
    pack = mp.InstancePack(tr.data + te.data)
    eng = mo.RL_PSO_Engine(pack, np.repeat(np.arange(24), 64), 60000, 50, seed=3)
    eng.init_population()
    prev_p, prev_g = eng.pbest.clone(), eng.gbest_val.clone()
    for step in range(150):
        state, reward, done = eng.update(torch.rand(eng.B, device=eng.device))
    assert torch.all(eng.fes == 250) and torch.all(eng.cur_index == 50) and torch.all(eng.pbest <= prev_p)
    assert torch.equal(eng.gbest_val, torch.minimum(prev_g, eng.pbest.min(dim=1).values))
    assert torch.all(eng.x.abs() <= 5) and torch.all(eng.v.abs() <= 1) and torch.all(torch.isfinite(state))
    with pytest.raises(ValueError):
        eng.update(torch.zeros(3))
