"""GPU: parity AT THE BENCHMARKED SHAPES (BASELINE.json: dim 30, pop 100) through the step kernels' own evaluator call.

Every function of the benchmarked suites goes through the fused step kernel at exactly D = 30, NP = 100 with the production RNG
(Philox in-kernel), and every launch is replayed by the oracle (`oracle/`, pinned against the unmodified reference by
tests/test_oracle_golden.py) with the numpy Philox twin from the GPU's previous state:

  * RLEPSO (`rlepso_step_kernel<30,BBOB>`, rlepso_optimizer.py:173-263): all 24 BBOB functions, 12 generations, two of them with forced
    stochastic re-initialisation (the second evaluation pass);
  * LDE (`lde_step_kernel<30,BBOB>`, lde_optimizer.py:159-198): all 30 noisy functions at NP = 100 (config C4), 10 generations;
  * DE-DDQN (`deddqn_step_kernel<30,BBOB>`, de_ddqn_optimizer.py:131-220): the 18 functions of the bbob-difficult test set (config C3),
    210 consecutive one-evaluation steps (two generation roll-overs of the pointer);

plus a distribution-level check (SURVEY.md section 4 plan, src/tester.py:196-205): 51 complete Philox rollouts on the GPU against 51
complete oracle rollouts (np.random.seed(run) like the tester) with the shipped policy on 3 functions -- two-sample Kolmogorov-Smirnov
on the final cost and on the consumed evaluations, p > 0.01.

Positions / populations are compared bit for bit, float32 costs to 1e-6 relative (BASELINE.json north_star)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
DIM, NP, MAXFES = 30, 100, 60000


def _suite(suit, fids=None, difficulty='easy', subset='all', dim=DIM):
    from metabox_b200 import problem as mp
    from oracle import bbob as ob
    tr, te = mp.BBOB_Dataset.get_datasets(suit, dim, 5, difficulty=difficulty)
    probs = te.data if subset == 'test' else tr.data + te.data
    otr, ote = ob.make_suite(suit, dim, 5.0)
    oin = {i.fid: i for i in otr + ote}
    if fids is not None:
        probs = [p for p in probs if p.fid in fids]
    return probs, [oin[p.fid] for p in probs]


def _weights(group):
    G = np.load(os.path.join(GOLDEN, 'agents.npz'))
    return {k[len(group) + 1:]: G[k] for k in G.files if k.startswith(group + '/')}


def test_rlepso_all_24_functions_d30_np100_philox_replay():
    import torch
    from metabox_b200 import optimizer as mo, problem as mp
    from oracle import bbob as ob, rlepso as orl
    from test_gpu_rlepso import _check
    probs, oins = _suite('bbob')
    assert sorted(p.fid for p in probs) == list(range(1, 25))
    pack = mp.InstancePack(probs)
    seed, base = 0xB200D30, 17
    eng = mo.RLEPSO_Engine(pack, list(range(len(probs))), MAXFES, 50, seed=seed, lane_base=base)
    cfg = orl.Config(dim=DIM, maxFEs=MAXFES)
    assert cfg.NP == NP
    eng.init_population()
    for b, I in enumerate(oins):
        ev = lambda x, draws, I=I: ob.func_true(I, x)
        S = orl.init_population(cfg, ev, I.optimum, -5., 5., orl.PhiloxTape(NP, DIM, seed, base + b, 0xFFFFFFFF))
        _check(eng.read_state(b), S, f'init fid {I.fid}')
    g = np.random.RandomState(30)
    n_reinit = 0
    for gen in range(12):
        action = g.rand(len(probs), 35).astype(np.float32)
        if gen in (4, 9):               # stagnation counters up -> the re-initialisation mask fires (rlepso_optimizer.py:238, 134-168)
            eng.pni += 40
        prev = [eng.read_state(b) for b in range(len(probs))]
        eng.update(torch.as_tensor(action).to(eng.device))
        for b, I in enumerate(oins):
            ev = lambda x, draws, I=I: ob.func_true(I, x)
            p = prev[b]
            S = orl.State(x=p['x'], v=p['v'], pbest_x=p['pbest_x'], c_cost=p['c_cost'], pbest=p['pbest'], pni=p['pni'],
                          gbest_x=p['gbest_x'], gbest_val=np.float32(p['gbest_val']), gbest_idx=int(p['gbest_idx']),
                          no_improve=0, fes=int(p['fes']), log_index=int(p['log_index']), curve=list(p['curve'][:int(p['curve_len'])]))
            N, _, _, _ = orl.update(cfg, S, action[b], ev, I.optimum, -5., 5., orl.PhiloxTape(NP, DIM, seed, base + b, gen))
            st = eng.read_state(b)
            _check(st, N, f'gen {gen} fid {I.fid}')
            n_reinit += int(int(st['fes']) - int(p['fes']) > NP)
    assert n_reinit >= 24, n_reinit     # every function went through the second evaluation pass at least once on average


def test_lde_all_30_noisy_functions_d30_np100_philox_replay():
    import torch
    from metabox_b200 import optimizer as mo, problem as mp
    from oracle import bbob as ob, lde as ol
    from test_gpu_lde import _check
    probs, oins = _suite('bbob-noisy')
    assert sorted(p.fid for p in probs) == list(range(101, 131))
    pack = mp.InstancePack(probs)
    seed, base = 0x1DE0D30, 5
    eng = mo.LDE_Engine(pack, list(range(len(probs))), MAXFES, 50, seed=seed, NP=NP, lane_base=base)
    cfg = ol.Config(dim=DIM, maxFEs=MAXFES, NP=NP)
    f = eng.init_population()
    for b, I in enumerate(oins):
        ev = lambda x, draws, I=I: ob.apply_noise(I, ob.func_true(I, x), draws=draws)
        S, feat = ol.init_population(cfg, ev, I.optimum, -5., 5., ol.PhiloxTape(NP, DIM, seed, base + b, 0xFFFFFFFF, I.noise))
        _check(eng.read_state(b), S, f[b].cpu().numpy(), feat, f'init fid {I.fid}')
    g = np.random.RandomState(31)
    for gen in range(10):
        action = g.rand(len(probs), 2 * NP).astype(np.float32)
        prev = [eng.read_state(b) for b in range(len(probs))]
        fg, rg, dg = eng.update(torch.as_tensor(action).to(eng.device))
        fg = fg.cpu().numpy()
        for b, I in enumerate(oins):
            ev = lambda x, draws, I=I: ob.apply_noise(I, ob.func_true(I, x), draws=draws)
            p = prev[b]
            S = ol.State(pop=p['pop'], fit=p['fit'], hist_sum=p['hist_sum'], hist_cnt=int(p['hist_cnt']), gbest=p['gbest'],
                         fes=int(p['fes']), log_index=int(p['log_index']), curve=list(p['curve'][:int(p['curve_len'])]))
            N, feat, reward, done = ol.update(cfg, S, action[b], ev, I.optimum, -5., 5., ol.PhiloxTape(NP, DIM, seed, base + b, gen, I.noise))
            _check(eng.read_state(b), N, fg[b], feat, f'gen {gen} fid {I.fid}')
            np.testing.assert_allclose(float(rg[b]), reward, rtol=2e-5, atol=1e-9)


def test_deddqn_18_difficult_functions_d30_np100_philox_replay():
    import torch
    from metabox_b200 import optimizer as mo, problem as mp
    from oracle import bbob as ob, deddqn as od
    from test_gpu_deddqn import _check
    probs, oins = _suite('bbob', difficulty='difficult', subset='test')
    assert len(probs) == 18
    pack = mp.InstancePack(probs)
    seed, base = 0xDD0D30, 2
    eng = mo.DE_DDQN_Engine(pack, list(range(len(probs))), MAXFES, 50, seed=seed, lane_base=base)
    cfg = od.Config(dim=DIM, maxFEs=MAXFES)
    assert cfg.NP == NP
    f = eng.init_population().cpu().numpy()
    states = []
    for b, I in enumerate(oins):
        ev = lambda x, draws, I=I: ob.apply_noise(I, ob.func_true(I, x), draws=draws)
        S, feat = od.init_population(cfg, ev, I.optimum, -5., 5., od.PhiloxTape(NP, DIM, seed, base + b, 0xFFFFFFFF, I.noise))
        _check(eng.read_state(b), S, f[b], feat, f'init fid {I.fid}')
        states.append(S)
    g = np.random.RandomState(32)
    for step in range(210):
        action = g.randint(0, 4, size=len(probs)).astype(np.int32)
        fg, rg, dg = eng.update(torch.as_tensor(action).to(eng.device))
        fg = fg.cpu().numpy()
        for b, I in enumerate(oins):
            ev = lambda x, draws, I=I: ob.apply_noise(I, ob.func_true(I, x), draws=draws)
            S = states[b]
            S, feat, reward, done = od.update(cfg, S, int(action[b]), ev, I.optimum, -5., 5., od.PhiloxTape(NP, DIM, seed, base + b, S.fes, I.noise))
            st = eng.read_state(b)
            _check(st, S, fg[b], feat, f'step {step} fid {I.fid}')
            # the oracle continues from the GPU's float32 costs so that 1e-7-level differences cannot accumulate over the 210 steps
            S.cost = st['cost'].copy(); S.c_gbest, S.c_gworst = np.float32(st['c3'][0]), np.float32(st['c3'][1])
            S.om_sum, S.om_max = st['om_sum'].copy(), st['om_max'].copy()
            w = st['win'][:S.win_len]; S.win[:S.win_len] = w[np.argsort(w[:, 6])][:, :6]
            states[b] = S


def _ks_p(a, b):
    from scipy import stats
    return float(stats.ks_2samp(a, b).pvalue)


def test_rollout_distributions_match_the_oracle_over_51_runs():
    """51 complete GPU rollouts (Philox, shipped RLEPSO policy) vs 51 complete oracle rollouts (MT19937 seeded per run like tester.py:198) on
    Sphere, Rastrigin_F15 and Schwefel at dim 10: the two samples of the final cost and of the consumed evaluations must be draws of the
    same distribution (two-sample KS, p > 0.01; both implementations are bit-compatible per step, so only the random streams differ)."""
    import torch
    from metabox_b200 import agent as ma, problem as mp, sweep
    from oracle import bbob as ob, rollout as orr
    dim, maxFEs, runs = 10, 20000, 51
    fids = [1, 15, 20]
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', dim, 5)
    probs = [[p for p in tr.data + te.data if p.fid == f][0] for f in fids]
    otr, ote = ob.make_suite('bbob', dim, 5.0)
    oins = [[i for i in otr + ote if i.fid == f][0] for f in fids]
    pack = mp.InstancePack(probs)
    inst, _ = sweep.lane_table(len(fids), runs)
    W = _weights('rlepso')
    torch.manual_seed(0)
    actor = ma.load_reference_weights(ma.RLEPSO_Actor(), W).cuda().eval()
    r = sweep.gather_results(sweep.RLEPSO_Sweep(pack, inst, actor, maxFEs, 50, seed=20261017).run(), 1).cpu().numpy().reshape(len(fids), runs, 53)
    wt = {k: torch.as_tensor(v) for k, v in W.items()}
    for k, I in enumerate(oins):
        o = [orr.rlepso_rollout(I, dim, maxFEs, wt, run) for run in range(runs)]
        o_cost = np.array([x['cost'][-1] for x in o]); o_fes = np.array([x['fes'] for x in o], dtype=np.float64)
        g_cost, g_fes = r[k, :, 50].astype(np.float64), r[k, :, 51].astype(np.float64)
        if np.all(o_cost <= 1e-8) and np.all(g_cost <= 1e-8):           # both always converge (Sphere): the cost carries no information
            p_cost = 1.0
        else:
            p_cost = _ks_p(np.log10(np.maximum(g_cost, 1e-12)), np.log10(np.maximum(o_cost, 1e-12)))
        p_fes = 1.0 if (np.ptp(o_fes) == 0 and np.ptp(g_fes) == 0 and o_fes[0] == g_fes[0]) else _ks_p(g_fes, o_fes)
        assert p_cost > 0.01 and p_fes > 0.01, (I.fid, p_cost, p_fes, np.median(g_cost), np.median(o_cost), g_fes.mean(), o_fes.mean())


def test_dim_20_the_third_compiled_dimension_through_all_step_kernels():
    """The reference takes any `--dim` (src/config.py:10); the library compiles a list of dimensions (csrc/mbx_launch.h MBX_FOR_BBOB_DIMS:
    10, 20, 30).  dim = 20 through the RLEPSO, LDE and DE-DDQN step kernels (Philox replay by the oracle, as above) and through the persistent
    DE-DDQN kernel and GLEET / Random_search (smoke: they run and keep their invariants)."""
    import torch
    from metabox_b200 import optimizer as mo, problem as mp
    from oracle import bbob as ob, deddqn as od, lde as ol, rlepso as orl
    from test_gpu_deddqn import _check as dd_check
    from test_gpu_lde import _check as lde_check
    from test_gpu_rlepso import _check as rl_check
    dim, maxFEs = 20, 40000
    probs, oins = _suite('bbob', fids=(1, 3, 7, 12, 16, 19, 21, 23), dim=dim)
    nprobs, noins = _suite('bbob-noisy', fids=(103, 113, 126), dim=dim)
    pack, npack = mp.InstancePack(probs), mp.InstancePack(nprobs)
    g = np.random.RandomState(20)
    # ---- RLEPSO
    seed = 0x200
    eng = mo.RLEPSO_Engine(pack, list(range(len(probs))), maxFEs, 50, seed=seed)
    cfg = orl.Config(dim=dim, maxFEs=maxFEs)
    eng.init_population()
    for gen in range(6):
        action = g.rand(len(probs), 35).astype(np.float32)
        if gen == 3:
            eng.pni += 40
        prev = [eng.read_state(b) for b in range(len(probs))]
        eng.update(torch.as_tensor(action).to(eng.device))
        for b, I in enumerate(oins):
            ev = lambda x, draws, I=I: ob.func_true(I, x)
            p = prev[b]
            S = orl.State(x=p['x'], v=p['v'], pbest_x=p['pbest_x'], c_cost=p['c_cost'], pbest=p['pbest'], pni=p['pni'],
                          gbest_x=p['gbest_x'], gbest_val=np.float32(p['gbest_val']), gbest_idx=int(p['gbest_idx']),
                          no_improve=0, fes=int(p['fes']), log_index=int(p['log_index']), curve=list(p['curve'][:int(p['curve_len'])]))
            N, _, _, _ = orl.update(cfg, S, action[b], ev, I.optimum, -5., 5., orl.PhiloxTape(cfg.NP, dim, seed, b, gen))
            rl_check(eng.read_state(b), N, f'rlepso d20 gen {gen} fid {I.fid}')
    # ---- LDE (noisy functions, NP 50)
    seed = 0x201
    eng = mo.LDE_Engine(npack, list(range(len(nprobs))), maxFEs, 50, seed=seed, NP=50)
    lcfg = ol.Config(dim=dim, maxFEs=maxFEs, NP=50)
    eng.init_population()
    for gen in range(5):
        action = g.rand(len(nprobs), 100).astype(np.float32)
        prev = [eng.read_state(b) for b in range(len(nprobs))]
        fg, rg, dg = eng.update(torch.as_tensor(action).to(eng.device))
        fg = fg.cpu().numpy()
        for b, I in enumerate(noins):
            ev = lambda x, draws, I=I: ob.apply_noise(I, ob.func_true(I, x), draws=draws)
            p = prev[b]
            S = ol.State(pop=p['pop'], fit=p['fit'], hist_sum=p['hist_sum'], hist_cnt=int(p['hist_cnt']), gbest=p['gbest'],
                         fes=int(p['fes']), log_index=int(p['log_index']), curve=list(p['curve'][:int(p['curve_len'])]))
            N, feat, reward, done = ol.update(lcfg, S, action[b], ev, I.optimum, -5., 5., ol.PhiloxTape(50, dim, seed, b, gen, I.noise))
            lde_check(eng.read_state(b), N, fg[b], feat, f'lde d20 gen {gen} fid {I.fid}')
    # ---- DE-DDQN
    seed = 0x202
    eng = mo.DE_DDQN_Engine(pack, list(range(len(probs))), maxFEs, 50, seed=seed)
    dcfg = od.Config(dim=dim, maxFEs=maxFEs)
    f = eng.init_population().cpu().numpy()
    states = []
    for b, I in enumerate(oins):
        ev = lambda x, draws, I=I: ob.apply_noise(I, ob.func_true(I, x), draws=draws)
        S, feat = od.init_population(dcfg, ev, I.optimum, -5., 5., od.PhiloxTape(dcfg.NP, dim, seed, b, 0xFFFFFFFF, I.noise))
        dd_check(eng.read_state(b), S, f[b], feat, f'deddqn d20 init fid {I.fid}')
        states.append(S)
    for step in range(40):
        action = g.randint(0, 4, size=len(probs)).astype(np.int32)
        fg, rg, dg = eng.update(torch.as_tensor(action).to(eng.device))
        fg = fg.cpu().numpy()
        for b, I in enumerate(oins):
            ev = lambda x, draws, I=I: ob.apply_noise(I, ob.func_true(I, x), draws=draws)
            S = states[b]
            S, feat, reward, done = od.update(dcfg, S, int(action[b]), ev, I.optimum, -5., 5., od.PhiloxTape(dcfg.NP, dim, seed, b, S.fes, I.noise))
            st = eng.read_state(b)
            dd_check(st, S, fg[b], feat, f'deddqn d20 step {step} fid {I.fid}')
            S.cost = st['cost'].copy(); S.c_gbest, S.c_gworst = np.float32(st['c3'][0]), np.float32(st['c3'][1])
            S.om_sum, S.om_max = st['om_sum'].copy(), st['om_max'].copy()
            w = st['win'][:S.win_len]; S.win[:S.win_len] = w[np.argsort(w[:, 6])][:, :6]
            states[b] = S
    # ---- GLEET and Random_search: run and keep their invariants
    ge = mo.GLEET_Engine(pack, list(range(len(probs))), maxFEs, 50, seed=5)
    ge.init_population()
    for _ in range(3):
        ge.update(torch.rand(len(probs), 100, device=ge.device))
    assert torch.all(ge.fes == 400) and torch.isfinite(ge.gbest_val).all() and torch.all(ge.gbest_val <= ge.pbest.min(dim=1).values + 0)
    import argparse
    rs = mo.Random_search(argparse.Namespace(dim=dim, maxFEs=2000, n_logpoint=50, log_interval=40)).run_lanes(pack, [0, 1], seed=3)
    assert all(r['fes'] == 2000 and all(a >= b for a, b in zip(r['cost'], r['cost'][1:])) for r in rs)
