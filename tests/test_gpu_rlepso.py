"""GPU: RLEPSO fused step parity -- CUDA (through the C ABI) vs golden reference trajectories, the oracle
(tape mode and Philox mode), plus size-independent invariants at the benchmark size."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
ATOL = 1e-10


def _problem(fid, dim):
    """-> (product problem, oracle eval closure, optimum or None, lb, ub, noise draw counts, oracle inst)"""
    from metabox_b200 import problem as mp
    from oracle import bbob as ob, protein as op
    if fid.isdigit():
        f = int(fid)
        suit = 'bbob' if f < 100 else 'bbob-noisy'
        tr, te = mp.BBOB_Dataset.get_datasets(suit, dim, 5)
        P = [p for p in tr.data + te.data if p.fid == f][0]
        otr, ote = ob.make_suite(suit, dim, 5.0)
        I = [i for i in otr + ote if i.fid == f][0]
        ev = lambda x, draws: ob.apply_noise(I, ob.func_true(I, x), draws=draws)
        return P, ev, I.optimum, -5.0, 5.0, ob.noise_draw_count(I), I
    G = np.load(os.path.join(GOLDEN, 'protein.npz'))
    a = [G[f'{k}_{fid}'] for k in ('coor_init', 'q', 'e', 'r', 'basis', 'eigval')]
    P = mp.Protein_Docking(*a, fid)
    I = op.ProteinInstance(fid, *a)
    return P, (lambda x, draws: op.energy(I, x)), None, op.LB, op.UB, (0, 0), I


def _tape_from_rec(rec, NP, D):
    """NumpyTape.rec (one lane) -> dict of arrays shaped for MbxRlepsoTape (B = 1)."""
    t = {}
    for k in ('init_pos_u', 'init_vel_u', 'clpso_u', 'fdr_u', 'reinit_pos_u', 'reinit_vel_u'):
        t[k] = rec.get(k, np.zeros((NP, D)))[None]
    t['tour'] = rec.get('tour', np.zeros((NP, D, 2), dtype=np.int64))[None].astype(np.int32)
    for k in ('rand1', 'rand2', 'reinit_u'):
        t[k] = rec.get(k, np.zeros(NP))[None]
    nz = np.zeros((1, 2, 3, NP))
    for w in (0, 1):
        for j, a in enumerate(rec.get(f'noise{w}') or ()):
            nz[0, w, j] = a
    t['noise'] = nz
    return t


def _check(st, S, tag, exact_cost=False):
    for k in ('x', 'v', 'pbest_x', 'gbest_x'):
        np.testing.assert_array_equal(st[k], getattr(S, k), err_msg=f'{tag} {k}')
    for k in ('c_cost', 'pbest'):
        np.testing.assert_allclose(st[k], getattr(S, k), rtol=1e-6, atol=ATOL, err_msg=f'{tag} {k}')
    np.testing.assert_array_equal(st['pni'], S.pni, err_msg=f'{tag} pni')
    assert int(st['gbest_idx']) == S.gbest_idx, tag
    assert int(st['fes']) == S.fes and int(st['log_index']) == S.log_index, tag
    np.testing.assert_allclose(st['curve'][:int(st['curve_len'])], np.array(S.curve, dtype=np.float64), rtol=1e-6, atol=ATOL)
    assert bool(st['done']) == bool(S.done), tag


def test_golden_trajectories_tape_mode():
    """Every reference RLEPSO step recorded in tests/golden/rlepso.npz (incl. re-initialisation, noisy and
    protein cases) is reproduced: positions/velocities/pbest bit-exact, FDR/tournament/argmin indices
    exact, costs within 1e-6 relative."""
    from metabox_b200 import optimizer as mo, problem as mp
    from oracle import rlepso as orl
    G = np.load(os.path.join(GOLDEN, 'rlepso.npz'))
    n_steps_total = n_reinit = 0
    for key in G['cases']:
        fid, dim = str(G[f'{key}/fid']), int(G[f'{key}/dim'])
        P, ev, optimum, lb, ub, nc, I = _problem(fid, dim)
        pack = mp.InstancePack([P])
        maxFEs, nlog = int(G[f'{key}/maxFEs']), int(G[f'{key}/n_logpoint'])
        cfg = orl.Config(dim=dim, maxFEs=maxFEs, n_logpoint=nlog)
        eng = mo.RLEPSO_Engine(pack, [0], maxFEs, nlog)
        # init
        np.random.seed(int(G[f'{key}/seed0']))
        tp = orl.NumpyTape(cfg.NP, dim, nc)
        S0 = orl.init_population(cfg, ev, optimum, lb, ub, tp)
        eng.init_population(tape=_tape_from_rec(tp.rec, cfg.NP, dim))
        st = eng.read_state(0)
        _check(st, S0, f'{key} init')
        for k in ('x', 'v', 'pbest_x'):                      # and against the reference's own arrays
            np.testing.assert_array_equal(st[k], G[f'{key}/s0/{k}'])
        for t in range(1, int(G[f'{key}/n_steps']) + 1):
            pre = f'{key}/s{t - 1}/'
            S = orl.State(x=G[pre + 'x'], v=G[pre + 'v'], pbest_x=G[pre + 'pbest_x'], c_cost=G[pre + 'c_cost'],
                          pbest=G[pre + 'pbest'], pni=G[f'{key}/s{t}/pni_in'], gbest_x=G[pre + 'gbest_x'],
                          gbest_val=np.float32(G[pre + 'gbest_val']), gbest_idx=int(G[pre + 'gbest_idx']), no_improve=0,
                          fes=int(G[pre + 'fes']), log_index=int(G[pre + 'log_index']),
                          curve=[np.float32(c) for c in G[pre + 'curve']])
            cur = np.zeros(51, dtype=np.float32); cur[:len(S.curve)] = S.curve
            eng.load_state(0, x=S.x, v=S.v, pbest_x=S.pbest_x, c_cost=S.c_cost, pbest=S.pbest, pni=S.pni,
                           gbest_x=S.gbest_x, gbest_val=S.gbest_val, gbest_idx=S.gbest_idx, fes=S.fes,
                           log_index=S.log_index, curve=cur, curve_len=len(S.curve), done=0)
            np.random.seed(int(G[f'{key}/s{t}/seed']))
            tp = orl.NumpyTape(cfg.NP, dim, nc)
            aux = {}
            action = G[f'{key}/s{t}/action']
            N, feat, reward, done = orl.update(cfg, S, action, ev, optimum, lb, ub, tp, aux)
            import torch
            dbg = torch.zeros(1, cfg.NP, dim, dtype=torch.int32, device=eng.device)
            f, r, d = eng.update(torch.as_tensor(action[None]).to(eng.device), tape=_tape_from_rec(tp.rec, cfg.NP, dim), dbg=dbg)
            st = eng.read_state(0)
            np.testing.assert_array_equal(dbg[0].cpu().numpy(), aux['fdr_target'], err_msg=f'{key} s{t} FDR targets')
            _check(st, N, f'{key} step {t}')
            post = f'{key}/s{t}/'
            for k in ('x', 'v', 'pbest_x', 'gbest_x'):      # the reference's own outputs, bit-exact
                np.testing.assert_array_equal(st[k], G[post + k], err_msg=f'{key} s{t} {k} vs reference')
            np.testing.assert_allclose(st['pbest'], G[post + 'pbest'], rtol=1e-6, atol=ATOL)
            assert int(st['gbest_idx']) == int(G[post + 'gbest_idx']) and int(st['fes']) == int(G[post + 'fes'])
            assert float(r[0]) == float(G[post + 'reward']) and bool(d[0]) == bool(G[post + 'done'])
            assert abs(float(f[0, 0]) - float(G[post + 'feature'][0])) < 1e-7
            n_steps_total += 1
            n_reinit += int(aux['reinit_mask'].any())
    assert n_steps_total >= 40 and n_reinit >= 5


def test_philox_production_path_vs_oracle():
    """Production RNG: each CUDA generation (Philox in-kernel) is replayed by the oracle with the numpy Philox
    twin from the GPU's previous state -- all state bit-exact over 40 generations, 8 lanes."""
    import torch
    from metabox_b200 import optimizer as mo, problem as mp
    from oracle import bbob as ob, rlepso as orl
    dim, maxFEs = 10, 20000
    fids = [1, 3, 6, 9, 12, 17, 23, 24]
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', dim, 5)
    probs = {p.fid: p for p in tr.data + te.data}
    otr, ote = ob.make_suite('bbob', dim, 5.0)
    oin = {i.fid: i for i in otr + ote}
    pack = mp.InstancePack([probs[f] for f in fids])
    seed = 0xC0FFEE1234
    eng = mo.RLEPSO_Engine(pack, list(range(len(fids))), maxFEs, 50, seed=seed, lane_base=5)
    cfg = orl.Config(dim=dim, maxFEs=maxFEs)
    eng.init_population()
    states = []
    for b, f in enumerate(fids):
        I = oin[f]
        ev = lambda x, draws, I=I: ob.func_true(I, x)
        S = orl.init_population(cfg, ev, I.optimum, -5., 5., orl.PhiloxTape(cfg.NP, dim, seed, 5 + b, 0xFFFFFFFF))
        _check(eng.read_state(b), S, f'init lane {b}')
        states.append(S)
    g = np.random.RandomState(3)
    mism = 0
    for gen in range(40):
        action = g.rand(len(fids), 35).astype(np.float32)
        if gen % 7 == 3:            # push per_no_improve up so that re-initialisation happens in Philox mode too
            eng.pni += 30
        prev = [eng.read_state(b) for b in range(len(fids))]
        eng.update(torch.as_tensor(action).to(eng.device))
        for b, f in enumerate(fids):
            I = oin[f]
            ev = lambda x, draws, I=I: ob.func_true(I, x)
            p = prev[b]
            S = orl.State(x=p['x'], v=p['v'], pbest_x=p['pbest_x'], c_cost=p['c_cost'], pbest=p['pbest'], pni=p['pni'],
                          gbest_x=p['gbest_x'], gbest_val=np.float32(p['gbest_val']), gbest_idx=int(p['gbest_idx']),
                          no_improve=0, fes=int(p['fes']), log_index=int(p['log_index']),
                          curve=list(p['curve'][:int(p['curve_len'])]))
            N, _, _, _ = orl.update(cfg, S, action[b], ev, I.optimum, -5., 5., orl.PhiloxTape(cfg.NP, dim, seed, 5 + b, gen))
            _check(eng.read_state(b), N, f'gen {gen} lane {b} (fid {f})')


def test_dropin_single_problem_api():
    """`PBO_Env(problem, RLEPSO_Optimizer(config))` used exactly like the reference's rollout loop."""
    import argparse
    from metabox_b200 import optimizer as mo, problem as mp, environment as me
    cfg = argparse.Namespace(dim=10, maxFEs=20000, n_logpoint=50, log_interval=400)
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', 10, 5)
    problem = te.data[0]
    env = me.PBO_Env(problem, mo.RLEPSO_Optimizer(cfg))
    np.random.seed(0)
    state = env.reset()
    assert state.shape == (1,) and state.dtype == np.float64 and abs(state[0] - 100 / 20000) < 1e-9
    g = np.random.RandomState(0)
    done, R, n = False, 0, 0
    while not done:
        state, reward, done = env.step(g.rand(35).astype(np.float32))
        assert reward in (1, -1) and isinstance(done, bool)
        R += reward; n += 1
    opt = env.optimizer
    assert opt.fes >= 20000 or opt.cost[-1] <= 1e-8
    assert 2 <= len(opt.cost) <= 51 and all(a >= b for a, b in zip(opt.cost, opt.cost[1:]))
    with pytest.raises(AssertionError):
        env.step(np.zeros(34, dtype=np.float32))


def test_invariants_at_benchmark_size():
    """BBOB dim=30 pop=100, 24 functions x 16 runs: size-independent properties of a PSO generation."""
    import torch
    from metabox_b200 import optimizer as mo, problem as mp
    dim, maxFEs, runs = 30, 60000, 16
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', dim, 5)
    probs = tr.data + te.data
    pack = mp.InstancePack(probs)
    lane_inst = np.repeat(np.arange(len(probs)), runs)
    eng = mo.RLEPSO_Engine(pack, lane_inst, maxFEs, 50, seed=11)
    eng.init_population()
    assert torch.all(eng.fes == 100) and torch.all(eng.curve_len == 1)
    assert torch.equal(eng.gbest_val, eng.pbest.min(dim=1).values)
    g = torch.Generator(device='cpu').manual_seed(0)
    prev_pbest, prev_gbest, prev_fes = eng.pbest.clone(), eng.gbest_val.clone(), eng.fes.clone()
    for gen in range(25):
        action = torch.rand(eng.B, 35, generator=g).to(eng.device)
        eng.update(action)
        assert torch.all(eng.x.abs() <= 5.0) and torch.all(eng.v.abs() <= 1.0)
        assert torch.all(eng.pbest <= prev_pbest)                           # pbest never worsens
        assert torch.equal(eng.gbest_val, torch.minimum(prev_gbest, eng.pbest.min(dim=1).values))
        bi = eng.gbest_idx.long()
        assert torch.all(eng.fes - prev_fes >= 100) and torch.all(eng.fes - prev_fes <= 200)
        assert torch.all(torch.isfinite(eng.c_cost))
        prev_pbest, prev_gbest, prev_fes = eng.pbest.clone(), eng.gbest_val.clone(), eng.fes.clone()
    # same (instance, seed, lane id) => same trajectory irrespective of batch composition / placement
    sub = [3, 40, 200, 383]
    eng2 = mo.RLEPSO_Engine(pack, lane_inst[sub[1]:sub[1] + 1], maxFEs, 50, seed=11, lane_base=sub[1])
    eng2.init_population()
    g = torch.Generator(device='cpu').manual_seed(0)
    for gen in range(25):
        action = torch.rand(eng.B, 35, generator=g)
        eng2.update(action[sub[1]:sub[1] + 1].to(eng.device))
    assert torch.equal(eng2.x[0], eng.x[sub[1]]) and torch.equal(eng2.pbest[0], eng.pbest[sub[1]])
    assert int(eng2.fes[0]) == int(eng.fes[sub[1]])


def test_done_lanes_are_frozen_and_empty_batch():
    import torch
    from metabox_b200 import optimizer as mo, problem as mp
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', 10, 5)
    pack = mp.InstancePack(te.data)
    eng = mo.RLEPSO_Engine(pack, [0, 0, 1], 300, 50, log_interval=6, seed=2)      # 300 FEs: done after 2 steps
    eng.init_population()
    a = torch.rand(3, 35).to(eng.device)
    eng.update(a); eng.update(a)
    assert torch.all(eng.done == 1)
    snap = (eng.x.clone(), eng.fes.clone(), eng.curve.clone())
    eng.update(a)
    assert torch.equal(snap[0], eng.x) and torch.equal(snap[1], eng.fes) and torch.equal(snap[2], eng.curve)
    e0 = mo.RLEPSO_Engine(pack, [], 300, 50, log_interval=6)
    e0.init_population(); e0.update(torch.zeros(0, 35).to(eng.device))


@pytest.mark.parametrize('dim', [10, 20, 30])
def test_fdr_targets_exact_incl_ties_duplicates_and_converged_swarms(dim):
    """FDR target indices (rlepso_optimizer.py:97-103) must equal np.argmin of the float64 quotients exactly.  The kernel
    screens in float32 and re-does ambiguous items with the reference's float64 divisions: crafted swarms hit both
    paths -- random, duplicated rows, all-equal costs, equal-cost groups, collapsed swarms (coordinates equal up to a few
    float32 ulps), huge / tiny cost scales."""
    import torch
    from metabox_b200 import optimizer as mo, problem as mp
    from oracle import rlepso as orl
    NP = 100
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', dim, 5)
    pack = mp.InstancePack(te.data[:1])
    g = np.random.RandomState(5)
    cases = []
    px = g.uniform(-5, 5, (NP, dim)); pf = g.uniform(0, 100, NP); cases.append(('random', px, pf))
    px2 = px.copy(); px2[10:40] = px2[0:30]; cases.append(('duplicated rows', px2, pf))
    cases.append(('all costs equal', px, np.full(NP, 3.25)))
    pf3 = np.repeat(g.uniform(0, 10, 10), 10); g.shuffle(pf3); cases.append(('equal-cost groups', px, pf3))
    c = g.uniform(-4, 4, dim).astype(np.float32).astype(np.float64)
    px4 = c[None, :] * (1 + g.randint(-3, 4, (NP, dim)) * 2.0 ** -23); cases.append(('collapsed swarm', px4, g.uniform(1e-8, 2e-8, NP)))
    px5 = c[None, :] + np.zeros((NP, dim)); cases.append(('identical positions', px5, g.uniform(0, 1, NP)))
    cases.append(('huge costs', px, g.uniform(0, 1, NP) * 1e30))
    cases.append(('tiny costs', px, g.uniform(0, 1, NP) * 1e-30))
    pf8 = pf.copy(); pf8[0:99:3] = pf8[1:100:3]; px8 = np.round(px * 4) / 4; cases.append(('grid positions + cost ties', px8, pf8))
    # near ties: equal distances (grid) and costs a few float32 ulps apart -> quotients 1e-7 .. 1e-5 apart, inside the resolution of the
    # screening keys (7 mantissa bits carry the candidate's position) but distinct in float64: every such item must go through pass 2
    base9 = np.float32(40.0); ulp = np.spacing(base9)
    pf9 = (base9 + ulp * g.randint(0, 400, NP)).astype(np.float32).astype(np.float64); pf9[::7] = g.uniform(50, 60, len(pf9[::7]))
    cases.append(('grid positions + costs ulps apart', np.round(px * 2) / 2, pf9))
    # the same on a continuous cloud: quotients of different candidates within 1e-5 relative by construction of the costs
    px10 = g.uniform(-5, 5, (NP, dim)); r10 = np.abs(px10[:, :1] - px10[0, 0]) + 1e-5
    pf10 = 100.0 - r10[:, 0] * (1.0 + 1e-5 * g.uniform(-1, 1, NP)); pf10[0] = 100.0; cases.append(('quotients within 1e-5', px10, pf10))
    cases.append(('costs of both signs, spread over 12 decades', px, np.sign(g.uniform(-1, 1, NP)) * 10.0 ** g.uniform(-6, 6, NP)))
    cases.append(('costs up to the float32 maximum', px, g.uniform(0, 1, NP) * 3e38))
    B = len(cases)
    eng = mo.RLEPSO_Engine(pack, [0] * B, 2000 * dim, 50, seed=1)
    eng.init_population()
    PX = np.stack([c_[1] for c_ in cases]).astype(np.float32); PF = np.stack([c_[2] for c_ in cases]).astype(np.float32)
    eng.pbest_x.copy_(torch.as_tensor(PX)); eng.pbest.copy_(torch.as_tensor(PF))
    dbg = torch.full((B, NP, dim), -1, dtype=torch.int32, device=eng.device)
    eng.update(torch.rand(B, 35).to(eng.device), dbg=dbg)
    got = dbg.cpu().numpy()
    for b, (name, _, _) in enumerate(cases):
        want = orl.fdr_targets(PX[b].astype(np.float64), PF[b].astype(np.float64))
        np.testing.assert_array_equal(got[b], want, err_msg=f'FDR targets: {name} (dim {dim})')


@pytest.mark.parametrize('dim', [10, 30])
def test_product_and_test_instantiations_of_the_step_kernel_agree_bit_for_bit(dim):
    """The step kernel is compiled twice (csrc/mbx_rlepso.cu): TEST = true accepts a tape / writes the FDR targets out and is what the
    oracle parity tests above drive; TEST = false is what every sweep launches.  Same source otherwise -- on Philox draws the two must leave
    identical bits in every state array, including re-initialisations, noise and lanes that finish."""
    import torch
    from metabox_b200 import optimizer as mo, problem as mp
    probs = []
    for suit, fids in (('bbob', (1, 7, 16, 21, 24)), ('bbob-noisy', (104, 118, 129))):
        tr, te = mp.BBOB_Dataset.get_datasets(suit, dim, 5)
        probs += [[p for p in tr.data + te.data if p.fid == f][0] for f in fids]
    pack = mp.InstancePack(probs)
    lanes = list(range(len(probs))) * 3
    A = mo.RLEPSO_Engine(pack, lanes, 4000, 50, seed=0xA11CE)
    Bn = mo.RLEPSO_Engine(pack, lanes, 4000, 50, seed=0xA11CE)
    A.init_population(); Bn.init_population()
    dbg = torch.full((A.B, A.NP, dim), -1, dtype=torch.int32, device=A.device)
    g = torch.Generator(device='cpu').manual_seed(3)
    for gen in range(42):                                   # 4000 FEs: every lane finishes inside the loop
        action = torch.rand(A.B, 35, generator=g).to(A.device)
        A.update(action)                                    # TEST = false
        Bn.update(action, dbg=dbg)                          # TEST = true (debug buffer)
        for k in ('x', 'v', 'pbest_x', 'pbest', 'c_cost', 'pni', 'gbest_x', 'gbest_val', 'gbest_idx', 'no_improve', 'fes', 'curve',
                  'curve_len', 'done', 'feature', 'reward'):
            assert torch.equal(getattr(A, k), getattr(Bn, k)), f'{k} differs after generation {gen}'
    assert bool((A.done != 0).all()) and int(dbg.min()) >= 0
