"""GPU: LDE fused step parity -- CUDA (through the C ABI) vs golden reference trajectories (tape mode), the
oracle replaying Philox counters (production mode), the drop-in single-problem API, and size-independent
invariants at BBOB-noisy dim=30, pop=100."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
ATOL = 1e-10

from test_gpu_rlepso import _problem  # noqa: E402  (same problem factory)


def _tape(rec, NP, D):
    """oracle NumpyTape.rec (one lane) -> dict of arrays shaped for MbxLdeTape (B = 1)."""
    t = {}
    for k in ('init_u', 'cross_u'):
        t[k] = np.asarray(rec.get(k, np.zeros((NP, D))))[None]
    t['pbest_idx'] = np.asarray(rec.get('pbest_idx', np.zeros(NP)))[None].astype(np.int32)
    dn = np.asarray(rec.get('donors', np.zeros((NP, 2), dtype=np.int64)))
    t['r1'] = dn[None, :, 0].astype(np.int32); t['r2'] = dn[None, :, 1].astype(np.int32)
    t['jrand'] = np.asarray(rec.get('jrand', np.zeros(NP)))[None].astype(np.int32)
    nz = np.zeros((1, 3, NP))
    for j, a in enumerate(rec.get('noise0') or ()):
        nz[0, j] = a
    t['noise'] = nz
    return t


def _check(st, S, feat_gpu, feat, tag):
    np.testing.assert_array_equal(st['pop'], S.pop, err_msg=f'{tag} pop (sorted)')
    np.testing.assert_allclose(st['fit'], S.fit, rtol=1e-6, atol=ATOL, err_msg=f'{tag} fit')
    np.testing.assert_array_equal(st['hist_sum'], S.hist_sum, err_msg=f'{tag} hist_sum')
    assert int(st['hist_cnt']) == S.hist_cnt and int(st['fes']) == S.fes and int(st['log_index']) == S.log_index, tag
    np.testing.assert_allclose(st['curve'][:int(st['curve_len'])], np.array(S.curve, dtype=np.float64), rtol=1e-6, atol=ATOL)
    assert bool(st['done']) == bool(S.done), tag
    NP = S.fit.shape[0]
    np.testing.assert_allclose(feat_gpu[:NP], feat[:NP], rtol=2e-5, atol=1e-7, err_msg=f'{tag} normalised fitness')
    np.testing.assert_array_equal(feat_gpu[NP:NP + 5], feat[NP:NP + 5], err_msg=f'{tag} histogram')
    np.testing.assert_allclose(feat_gpu[NP + 5:], feat[NP + 5:], rtol=1e-6, err_msg=f'{tag} mean past histogram')


def _load(eng, S):
    cur = np.zeros(51, dtype=np.float32); cur[:len(S.curve)] = S.curve
    eng.load_state(0, pop=S.pop, fit=S.fit, hist_sum=S.hist_sum, hist_cnt=S.hist_cnt, gbest=S.gbest, fes=S.fes,
                   log_index=S.log_index, curve=cur, curve_len=len(S.curve), done=0)


def test_golden_trajectories_tape_mode():
    """Every reference LDE step in tests/golden/lde.npz: trial vectors and the sorted survivor population bit-exact,
    `<=` selection / sort order / 5-bin histograms exact, costs within 1e-6 relative."""
    import torch
    from metabox_b200 import optimizer as mo, problem as mp
    from oracle import lde as ol
    G = np.load(os.path.join(GOLDEN, 'lde.npz'))
    n_steps = n_done = 0
    for key in G['cases']:
        fid, dim = str(G[f'{key}/fid']), int(G[f'{key}/dim'])
        P, ev, optimum, lb, ub, nc, I = _problem(fid, dim)
        pack = mp.InstancePack([P])
        maxFEs, nlog = int(G[f'{key}/maxFEs']), int(G[f'{key}/n_logpoint'])
        cfg = ol.Config(dim=dim, maxFEs=maxFEs, n_logpoint=nlog)
        eng = mo.LDE_Engine(pack, [0], maxFEs, nlog)
        s0 = int(G[f'{key}/seed0'])
        np.random.seed(s0); torch.manual_seed(s0)
        tp = ol.NumpyTape(cfg.NP, dim, nc)
        S0, feat0 = ol.init_population(cfg, ev, optimum, lb, ub, tp)
        f = eng.init_population(tape=_tape(tp.rec, cfg.NP, dim))
        _check(eng.read_state(0), S0, f[0].cpu().numpy(), feat0, f'{key} init')
        np.testing.assert_array_equal(eng.read_state(0)['pop'], G[f'{key}/s0/pop'])     # the reference's own array
        for t in range(1, int(G[f'{key}/n_steps']) + 1):
            pre, post = f'{key}/s{t - 1}/', f'{key}/s{t}/'
            ph = G[pre + 'past_histo']
            S = ol.State(pop=G[pre + 'pop'], fit=G[pre + 'fit'], hist_sum=ph.sum(0), hist_cnt=ph.shape[0],
                         gbest=np.float32(G[pre + 'gbest']), fes=int(G[pre + 'fes']), log_index=int(G[pre + 'log_index']),
                         curve=[np.float32(c) for c in G[pre + 'curve']])
            _load(eng, S)
            s = int(G[post + 'seed'])
            np.random.seed(s); torch.manual_seed(s)
            tp = ol.NumpyTape(cfg.NP, dim, nc)
            aux = {}
            action = G[post + 'action']
            N, feat, reward, done = ol.update(cfg, S, action, ev, optimum, lb, ub, tp, aux)
            dbg = torch.zeros(1, cfg.NP * (dim + 1), dtype=torch.float32, device=eng.device)
            fg, rg, dg = eng.update(torch.as_tensor(action[None]).to(eng.device), tape=_tape(tp.rec, cfg.NP, dim), dbg=dbg)
            d = dbg[0].cpu().numpy()
            np.testing.assert_array_equal(d[:cfg.NP * dim].reshape(cfg.NP, dim), aux['trial'], err_msg=f'{key} s{t} trial vectors')
            tf = d[cfg.NP * dim:]
            np.testing.assert_allclose(tf, aux['trial_fit'], rtol=1e-6, atol=ATOL)
            np.testing.assert_array_equal(tf <= S.fit, aux['surv'], err_msg=f'{key} s{t} survivor mask')
            st = eng.read_state(0)
            _check(st, N, fg[0].cpu().numpy(), feat, f'{key} step {t}')
            np.testing.assert_array_equal(st['pop'], G[post + 'pop'], err_msg=f'{key} s{t} pop vs reference')
            np.testing.assert_array_equal(fg[0].cpu().numpy()[cfg.NP:cfg.NP + 5], G[post + 'feature'][cfg.NP:cfg.NP + 5])
            np.testing.assert_allclose(float(rg[0]), float(G[post + 'reward']), rtol=2e-5, atol=1e-9)
            assert bool(dg[0]) == bool(G[post + 'done']) and int(st['fes']) == int(G[post + 'fes'])
            n_steps += 1; n_done += int(done)
    assert n_steps >= 30 and n_done >= 1


@pytest.mark.parametrize('NP', [50, 100])
def test_philox_production_path_vs_oracle(NP):
    """Production RNG (Philox in-kernel, rejection-free donors) replayed generation by generation by the oracle's
    numpy Philox twin from the GPU's previous state: populations bit-exact over 30 generations, incl. noisy ids."""
    import torch
    from metabox_b200 import optimizer as mo, problem as mp
    from oracle import bbob as ob, lde as ol
    dim, maxFEs = 10, 20000
    cases = [('bbob', 1), ('bbob', 8), ('bbob', 15), ('bbob', 20), ('bbob-noisy', 102), ('bbob-noisy', 121)]
    probs, oins = [], []
    for suit, f in cases:
        tr, te = mp.BBOB_Dataset.get_datasets(suit, dim, 5)
        probs.append([p for p in tr.data + te.data if p.fid == f][0])
        otr, ote = ob.make_suite(suit, dim, 5.0)
        oins.append([i for i in otr + ote if i.fid == f][0])
    pack = mp.InstancePack(probs)
    seed = 0xBADC0DE77
    eng = mo.LDE_Engine(pack, list(range(len(cases))), maxFEs, 50, seed=seed, NP=NP, lane_base=3)
    cfg = ol.Config(dim=dim, maxFEs=maxFEs, NP=NP)
    f = eng.init_population()
    for b, I in enumerate(oins):
        ev = lambda x, draws, I=I: ob.apply_noise(I, ob.func_true(I, x), draws=draws)
        S, feat = ol.init_population(cfg, ev, I.optimum, -5., 5., ol.PhiloxTape(NP, dim, seed, 3 + b, 0xFFFFFFFF, I.noise))
        _check(eng.read_state(b), S, f[b].cpu().numpy(), feat, f'init lane {b}')
    g = np.random.RandomState(5)
    for gen in range(30):
        action = g.rand(len(cases), 2 * NP).astype(np.float32)
        prev = [eng.read_state(b) for b in range(len(cases))]
        fg, rg, dg = eng.update(torch.as_tensor(action).to(eng.device))
        fg = fg.cpu().numpy()
        for b, I in enumerate(oins):
            ev = lambda x, draws, I=I: ob.apply_noise(I, ob.func_true(I, x), draws=draws)
            p = prev[b]
            S = ol.State(pop=p['pop'], fit=p['fit'], hist_sum=p['hist_sum'], hist_cnt=int(p['hist_cnt']), gbest=p['gbest'],
                         fes=int(p['fes']), log_index=int(p['log_index']), curve=list(p['curve'][:int(p['curve_len'])]))
            N, feat, reward, done = ol.update(cfg, S, action[b], ev, I.optimum, -5., 5.,
                                              ol.PhiloxTape(NP, dim, seed, 3 + b, gen, I.noise))
            _check(eng.read_state(b), N, fg[b], feat, f'gen {gen} lane {b} (fid {I.fid})')
            np.testing.assert_allclose(float(rg[b]), reward, rtol=2e-5, atol=1e-9)


def test_dropin_single_problem_api():
    """`PBO_Env(problem, LDE_Optimizer(config))` driven like `LDE_Agent.rollout_episode` (lde_agent.py:147-163)."""
    import argparse
    from metabox_b200 import optimizer as mo, problem as mp, environment as me
    cfg = argparse.Namespace(dim=10, maxFEs=5000, n_logpoint=50, log_interval=100)
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', 10, 5)
    env = me.PBO_Env(te.data[0], mo.LDE_Optimizer(cfg))
    np.random.seed(0)
    state = env.reset()
    assert state.shape == (1, 60) and state.dtype == np.float64
    assert state[0, 0] == 0. and abs(state[0, 49] - 1.) < 1e-7 and state[0, 50:55].sum() == 50 and np.all(state[0, 55:] == 10.)
    g = np.random.RandomState(0)
    done, n = False, 0
    while not done:
        state, reward, done = env.step(g.rand(1, 100).astype(np.float32))
        assert reward.shape == (1,) and reward[0] >= 0. and isinstance(done, bool)
        assert np.all(np.diff(state[0, :50]) >= 0)                      # sorted normalised fitness
        n += 1
    opt = env.optimizer
    assert opt.fes == 50 * (n + 1) and (opt.fes >= 5000 or opt.cost[-1] <= 1e-8)
    assert all(a >= b for a, b in zip(opt.cost, opt.cost[1:])) and opt.get_best() == opt.cost[-1]
    with pytest.raises(ValueError):
        env.step(np.zeros((1, 99), dtype=np.float32))


def test_invariants_at_benchmark_size():
    """bbob-noisy 30 functions, dim=30, pop=100, 8 runs each (BASELINE config C4 shape): DE generation properties."""
    import torch
    from metabox_b200 import optimizer as mo, problem as mp
    dim, maxFEs, runs, NP = 30, 60000, 8, 100
    tr, te = mp.BBOB_Dataset.get_datasets('bbob-noisy', dim, 5)
    probs = tr.data + te.data
    assert len(probs) == 30
    pack = mp.InstancePack(probs)
    lane_inst = np.repeat(np.arange(len(probs)), runs)
    eng = mo.LDE_Engine(pack, lane_inst, maxFEs, 50, seed=7, NP=NP)
    f = eng.init_population()
    assert torch.all(eng.fes == NP) and torch.all(eng.curve_len == 1) and torch.all(eng.hist_cnt == 1)
    g = torch.Generator(device='cpu').manual_seed(0)
    prev_fit = eng.fit.clone()
    for gen in range(20):
        action = torch.rand(eng.B, 2 * NP, generator=g).to(eng.device)
        f, r, d = eng.update(action)
        assert torch.all(eng.pop.abs() <= 5.0)
        assert torch.all(eng.fit[:, 1:] >= eng.fit[:, :-1])                       # sorted ascending
        assert torch.all(eng.fit.sum(1) <= prev_fit.sum(1) * (1 + 1e-6))           # `<=` selection never worsens anyone
        assert torch.equal(eng.gbest, eng.fit[:, 0])
        assert torch.all(f[:, NP:NP + 5].sum(1) == NP) and torch.all(f[:, 0] == 0)
        assert torch.allclose(f[:, NP + 5:].sum(1), torch.full((eng.B,), float(NP), device=eng.device))
        assert torch.all(eng.fes == NP * (gen + 2)) and torch.all(eng.hist_cnt == gen + 2)
        assert torch.all(r >= 0) and torch.all(torch.isfinite(eng.fit))
        prev_fit = eng.fit.clone()
    # placement independence: lane 77 alone reproduces lane 77 of the batch
    eng2 = mo.LDE_Engine(pack, lane_inst[77:78], maxFEs, 50, seed=7, NP=NP, lane_base=77)
    eng2.init_population()
    g = torch.Generator(device='cpu').manual_seed(0)
    for gen in range(20):
        action = torch.rand(eng.B, 2 * NP, generator=g)
        eng2.update(action[77:78].to(eng.device))
    assert torch.equal(eng2.pop[0], eng.pop[77]) and torch.equal(eng2.fit[0], eng.fit[77])


def test_done_lanes_are_frozen_and_empty_batch():
    import torch
    from metabox_b200 import optimizer as mo, problem as mp
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', 10, 5)
    pack = mp.InstancePack(te.data)
    eng = mo.LDE_Engine(pack, [0, 0, 1], 150, 50, log_interval=3, seed=2)          # 150 FEs: done after 2 steps
    eng.init_population()
    a = torch.rand(3, 100).to(eng.device)
    eng.update(a); eng.update(a)
    assert torch.all(eng.done == 1) and torch.all(eng.fes == 150)
    snap = (eng.pop.clone(), eng.fes.clone(), eng.curve.clone(), eng.hist_sum.clone())
    eng.update(a)
    assert all(torch.equal(s, t) for s, t in zip(snap, (eng.pop, eng.fes, eng.curve, eng.hist_sum)))
    e0 = mo.LDE_Engine(pack, [], 150, 50, log_interval=3)
    e0.init_population(); e0.update(torch.zeros(0, 100).to(eng.device))
