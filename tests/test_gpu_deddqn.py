"""GPU: DE-DDQN fused step parity -- CUDA (through the C ABI) teacher-forced from the oracle's state at every step
of the golden trajectories (the oracle itself is pinned to the reference by tests/test_oracle_golden.py), the
Philox production path replayed by the oracle, the drop-in API, and invariants at dim=30."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
ATOL = 1e-10

from test_gpu_rlepso import _problem  # noqa: E402


def _load(eng, S, b=0):
    """oracle State -> engine lane b."""
    cur = np.zeros(51, dtype=np.float32); cur[:len(S.curve)] = S.curve
    win = np.zeros((eng.W, eng.WS)); win[:, :6] = S.win; win[:, 6] = np.arange(eng.W)
    ints = [S.pointer, S.gen, S.stagcount, S.prebest_row, int(S.gbest_rebound), S.win_len, S.win_len, 0]
    eng.load_state(b, X=S.X, cost=S.cost, X_gbest=S.X_gbest, X_prebest=S.X_prebest, c3=[S.c_gbest, S.c_gworst, S.c_prebest],
                   ints=ints, r=S.r, n_tot=S.n_tot, n_succ=S.n_succ, om_sum=S.om_sum, om_max=S.om_max, win=win, fes=S.fes,
                   log_index=S.log_index, curve=cur, curve_len=len(S.curve), done=0)


def _check(st, S, feat_gpu, feat, tag):
    np.testing.assert_array_equal(st['X'], S.X, err_msg=f'{tag} X')
    np.testing.assert_allclose(st['cost'], S.cost, rtol=1e-6, atol=ATOL, err_msg=f'{tag} cost')
    np.testing.assert_array_equal(st['X_gbest'], S.X_gbest, err_msg=f'{tag} X_gbest')
    ints = st['ints']
    assert list(ints[:3]) == [S.pointer, S.gen, S.stagcount], f'{tag} pointer/gen/stagcount {ints[:3]}'
    assert int(ints[3]) == S.prebest_row and bool(ints[4]) == bool(S.gbest_rebound) and int(ints[5]) == S.win_len, tag
    if S.prebest_row < 0:
        np.testing.assert_array_equal(st['X_prebest'], S.X_prebest, err_msg=f'{tag} X_prebest')
    np.testing.assert_allclose(st['c3'], [S.c_gbest, S.c_gworst, S.c_prebest], rtol=1e-6, atol=ATOL, err_msg=f'{tag} c3')
    np.testing.assert_array_equal(st['r'], S.r, err_msg=f'{tag} r')
    np.testing.assert_array_equal(st['n_tot'], S.n_tot, err_msg=f'{tag} N_tot')
    np.testing.assert_array_equal(st['n_succ'], S.n_succ, err_msg=f'{tag} N_succ')
    np.testing.assert_allclose(st['om_sum'], S.om_sum, rtol=1e-5, atol=1e-12, err_msg=f'{tag} OM sums')
    np.testing.assert_allclose(st['om_max'], S.om_max, rtol=1e-5, atol=1e-12, err_msg=f'{tag} OM max')
    w = st['win'][:S.win_len]
    w = w[np.argsort(w[:, 6])]                                           # back to list order
    np.testing.assert_array_equal(w[:, 0], S.win[:S.win_len, 0], err_msg=f'{tag} window operators')
    np.testing.assert_allclose(w[:, 1:6], S.win[:S.win_len, 1:6], rtol=1e-5, atol=1e-12, err_msg=f'{tag} window')
    assert int(st['fes']) == S.fes and int(st['log_index']) == S.log_index and bool(st['done']) == bool(S.done), tag
    np.testing.assert_allclose(st['curve'][:int(st['curve_len'])], np.array(S.curve, dtype=np.float64), rtol=1e-6, atol=ATOL)
    with np.errstate(invalid='ignore'):
        np.testing.assert_allclose(feat_gpu, feat, rtol=3e-5, atol=2e-7, err_msg=f'{tag} features')


def test_golden_trajectories_tape_mode():
    """Along every golden DE_DDQN trajectory (reference actions / seeds): the CUDA step, started from the oracle's
    state, reproduces the trial vector bit-exactly, the replacement decision, all credit records and the 99 features."""
    import torch
    from metabox_b200 import optimizer as mo, problem as mp
    from oracle import deddqn as od
    G = np.load(os.path.join(GOLDEN, 'deddqn.npz'))
    n_steps = n_repl = 0
    for key in G['cases']:
        fid, dim = str(G[f'{key}/fid']), int(G[f'{key}/dim'])
        P, ev, optimum, lb, ub, nc, I = _problem(fid, dim)
        pack = mp.InstancePack([P])
        maxFEs, nlog = int(G[f'{key}/maxFEs']), int(G[f'{key}/n_logpoint'])
        cfg = od.Config(dim=dim, maxFEs=maxFEs, n_logpoint=nlog)
        eng = mo.DE_DDQN_Engine(pack, [0], maxFEs, nlog)
        s0 = int(G[f'{key}/seed0'])
        np.random.seed(s0)
        tp = od.NumpyTape(cfg.NP, dim, nc)
        S, feat = od.init_population(cfg, ev, optimum, lb, ub, tp)
        nz = np.zeros((1, 3, cfg.NP))
        for j, a in enumerate(tp.rec.get('noise') or ()):
            nz[0, j] = a
        f = eng.init_population(tape=dict(init_u=tp.rec['init_u'][None], r5=tp.rec['r'][None].astype(np.int32), noise=nz))
        _check(eng.read_state(0), S, f[0].cpu().numpy(), feat, f'{key} init')
        stride = 1 if int(G[f'{key}/n_steps']) <= 130 else 3                # check every step of the short cases
        for t in range(1, int(G[f'{key}/n_steps']) + 1):
            check = (t % stride == 0) or t <= 5 or (S.pointer in (0, 99))
            if check:
                _load(eng, S)
            np.random.seed(s0 + t)
            tp = od.NumpyTape(cfg.NP, dim, nc)
            aux = {}
            action = int(G[f'{key}/action'][t - 1])
            S, feat, reward, done = od.update(cfg, S, action, ev, optimum, lb, ub, tp, aux)
            if not check:
                continue
            nz = np.zeros((1, 3, 1))
            for j, a in enumerate(tp.rec.get('noise') or ()):
                nz[0, j] = a
            dbg = torch.zeros(1, dim + 1, dtype=torch.float32, device=eng.device)
            fg, rg, dg = eng.update(torch.tensor([action], dtype=torch.int32, device=eng.device),
                                    tape=dict(r5=tp.rec['r'][None].astype(np.int32), noise=nz), dbg=dbg)
            d = dbg[0].cpu().numpy()
            np.testing.assert_array_equal(d[:dim], aux['trial'], err_msg=f'{key} s{t} trial')
            np.testing.assert_allclose(d[dim], aux['trial_cost'], rtol=1e-6, atol=ATOL)
            _check(eng.read_state(0), S, fg[0].cpu().numpy(), feat, f'{key} step {t}')
            np.testing.assert_allclose(float(rg[0]), reward, rtol=1e-5, atol=1e-9)
            with np.errstate(invalid='ignore'):
                np.testing.assert_allclose(fg[0].cpu().numpy(), G[f'{key}/feature'][t], rtol=5e-5, atol=3e-7,
                                           err_msg=f'{key} s{t} features vs reference')
            assert bool(dg[0]) == bool(G[f'{key}/done'][t - 1])
            n_steps += 1; n_repl += int(aux['replaced'])
    assert n_steps >= 500 and n_repl >= 50


def test_philox_production_path_vs_oracle():
    """Production RNG: 260 consecutive CUDA steps (Philox in-kernel) on 5 lanes, each replayed by the oracle from the
    GPU's previous state."""
    import torch
    from metabox_b200 import optimizer as mo, problem as mp
    from oracle import bbob as ob, deddqn as od
    dim, maxFEs = 10, 20000
    cases = [('bbob', 1), ('bbob', 6), ('bbob', 16), ('bbob-noisy', 104), ('bbob-noisy', 124)]
    probs, oins = [], []
    for suit, f in cases:
        tr, te = mp.BBOB_Dataset.get_datasets(suit, dim, 5)
        probs.append([p for p in tr.data + te.data if p.fid == f][0])
        otr, ote = ob.make_suite(suit, dim, 5.0)
        oins.append([i for i in otr + ote if i.fid == f][0])
    pack = mp.InstancePack(probs)
    seed = 0x5EED5EED01
    eng = mo.DE_DDQN_Engine(pack, list(range(len(cases))), maxFEs, 50, seed=seed, lane_base=9)
    cfg = od.Config(dim=dim, maxFEs=maxFEs)
    f = eng.init_population().cpu().numpy()
    states = []
    for b, I in enumerate(oins):
        ev = lambda x, draws, I=I: ob.apply_noise(I, ob.func_true(I, x), draws=draws)
        S, feat = od.init_population(cfg, ev, I.optimum, -5., 5., od.PhiloxTape(cfg.NP, dim, seed, 9 + b, 0xFFFFFFFF, I.noise))
        _check(eng.read_state(b), S, f[b], feat, f'init lane {b}')
        states.append(S)
    g = np.random.RandomState(8)
    for step in range(260):
        action = g.randint(0, 4, size=len(cases)).astype(np.int32)
        fg, rg, dg = eng.update(torch.as_tensor(action).to(eng.device))
        fg = fg.cpu().numpy()
        for b, I in enumerate(oins):
            ev = lambda x, draws, I=I: ob.apply_noise(I, ob.func_true(I, x), draws=draws)
            S = states[b]
            fes0 = S.fes
            S, feat, reward, done = od.update(cfg, S, int(action[b]), ev, I.optimum, -5., 5.,
                                              od.PhiloxTape(cfg.NP, dim, seed, 9 + b, fes0, I.noise))
            st = eng.read_state(b)
            _check(st, S, fg[b], feat, f'step {step} lane {b} (fid {I.fid})')
            # keep the oracle on the GPU's float32 costs so that 1e-7-level differences cannot accumulate
            S.cost = st['cost'].copy(); S.c_gbest, S.c_gworst = np.float32(st['c3'][0]), np.float32(st['c3'][1])
            S.om_sum, S.om_max = st['om_sum'].copy(), st['om_max'].copy()
            w = st['win'][:S.win_len]; S.win[:S.win_len] = w[np.argsort(w[:, 6])][:, :6]
            states[b] = S


def test_dropin_single_problem_api():
    """`PBO_Env(problem, DE_DDQN_Optimizer(config))` driven like `DE_DDQN_Agent.rollout_episode` (de_ddqn_agent.py:108-117)
    with the shipped Q-network choosing greedy actions."""
    import argparse
    import torch
    from metabox_b200 import agent as ma, optimizer as mo, problem as mp, environment as me
    cfg = argparse.Namespace(dim=10, maxFEs=600, n_logpoint=50, log_interval=12)
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', 10, 5)
    env = me.PBO_Env(te.data[0], mo.DE_DDQN_Optimizer(cfg))
    net = ma.DE_DDQN_Net()
    G = np.load(os.path.join(GOLDEN, 'agents.npz'))
    ma.load_reference_weights(net, {k[len('deddqn/'):]: G[k] for k in G.files if k.startswith('deddqn/')})
    np.random.seed(0)
    state = env.reset()
    assert state.shape == (99,) and state.dtype == np.float64 and state[4] == 1.0 and abs(state[3] - 500 / 600) < 1e-6
    done, n, R = False, 0, 0.
    while not done:
        a = int(net.act(torch.as_tensor(state, dtype=torch.float32)))
        state, reward, done = env.step(a)
        assert reward >= 0 and isinstance(done, bool)
        R += reward; n += 1
    assert env.optimizer.fes == 100 + n and (env.optimizer.fes >= 600 or env.optimizer.cost[-1] <= 1e-8)
    assert all(x >= y for x, y in zip(env.optimizer.cost, env.optimizer.cost[1:]))
    with pytest.raises(ValueError):
        env.step(7)


def test_invariants_at_benchmark_size():
    """bbob difficult test set (18 fns), dim=30, pop=100, 4 runs each (BASELINE config C3 shape), 150 lockstep steps."""
    import torch
    from metabox_b200 import optimizer as mo, problem as mp
    dim, maxFEs, runs = 30, 60000, 4
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', dim, 5, difficulty='difficult')
    probs = te.data
    assert len(probs) == 18
    pack = mp.InstancePack(probs)
    lane_inst = np.repeat(np.arange(len(probs)), runs)
    eng = mo.DE_DDQN_Engine(pack, lane_inst, maxFEs, 50, seed=3)
    eng.init_population()
    assert torch.all(eng.fes == 100) and torch.equal(eng.c3[:, 0], eng.cost.min(1).values)
    g = torch.Generator(device='cpu').manual_seed(1)
    prev = eng.cost.clone()
    for step in range(150):
        a = torch.randint(0, 4, (eng.B,), generator=g).to(torch.int32).to(eng.device)
        f, r, d = eng.update(a)
        assert torch.all(eng.X.abs() <= 5.0)
        assert torch.all(eng.cost <= prev) and torch.all(r >= 0)
        assert torch.equal(eng.c3[:, 0], eng.cost.min(1).values)
        assert torch.all(eng.ints[:, 0] == (step + 1) % 100) and torch.all(eng.fes == 101 + step)
        assert torch.all(eng.ints[:, 5] == min(step + 1, 50)) and torch.all(eng.ints[:, 7] == 0)
        assert torch.all(eng.n_tot[:, :, 0].sum(1) == (step % 100) + 1)
        prev = eng.cost.clone()
    assert torch.all(eng.ints[:, 1] == 2)
    # an out-of-range action flags the lane and leaves it untouched
    snap = eng.fes.clone()
    a = torch.zeros(eng.B, dtype=torch.int32, device=eng.device); a[5] = 9
    eng.update(a)
    assert int(eng.ints[5, 7]) == 1 and int(eng.fes[5]) == int(snap[5]) and int(eng.fes[4]) == int(snap[4]) + 1
