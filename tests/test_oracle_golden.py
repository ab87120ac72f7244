"""CPU: the oracle restatement against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  This is what pins the oracle."""
import os

import numpy as np
import pytest

from oracle import bbob as ob
from oracle import protein as op
from oracle import rlepso as orl

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope='module', params=[10, 20, 30])
def suite(request):
    dim = request.param
    G = load(f'bbob_d{dim}.npz')
    insts = {}
    for suit in ('bbob', 'bbob-noisy'):
        tr, te = ob.make_suite(suit, dim, 5.0, 'easy')
        for I in tr + te:
            insts[I.fid] = I
    return dim, G, insts


def test_instance_generation_matches_reference(suite):
    dim, G, insts = suite
    assert sorted(insts) == list(range(1, 25)) + list(range(101, 131))
    for fid, I in insts.items():
        assert str(G[f'name_{fid}']) == I.name
        assert float(G[f'opt_{fid}']) == I.optimum          # bit-exact: same draws, same arithmetic
        if dim == 10:
            np.testing.assert_array_equal(G[f'shift_{fid}'], I.shift)
            np.testing.assert_array_equal(G[f'R_{fid}'], I.R)
            assert float(G[f'bias_{fid}']) == I.bias
            for k in ('M2', 'y', 'C', 'w'):
                if f'{k}_{fid}' in G.files:
                    np.testing.assert_array_equal(G[f'{k}_{fid}'], getattr(I, k))


def test_splits(suite):
    dim, G, _ = suite
    for suit in ('bbob', 'bbob-noisy'):
        for diff in ('easy', 'difficult'):
            _, te = ob.make_suite(suit, dim, 5.0, diff)
            assert [I.name for I in te] == list(G[f'split_{diff}_{suit}'])


def test_noise_free_fitness(suite):
    dim, G, insts = suite
    for fid, I in insts.items():
        X = G[f'x_{fid}'].astype(np.float64)
        F = ob.func_true(I, X)
        ref = G[f'f_{fid}']
        # same formulas, same libm: differences come only from summation order / array layout
        np.testing.assert_allclose(F, ref, rtol=2e-13, atol=0, err_msg=f'fid {fid}')
        cost, rcost = F - I.optimum, ref - float(G[f'opt_{fid}'])
        np.testing.assert_allclose(cost, rcost, rtol=1e-9, atol=64 * np.spacing(abs(I.bias) + 1.0))


def test_noisy_fitness(suite):
    dim, G, insts = suite
    for fid in range(101, 131):
        I = insts[fid]
        X = G[f'x_{fid}'].astype(np.float64)
        np.random.seed(777 + fid)
        F = ob.evaluate(I, X)
        np.testing.assert_allclose(F, G[f'fn_{fid}'], rtol=2e-13, err_msg=f'fid {fid}')
        np.random.seed(778 + fid)
        F1 = np.array([ob.evaluate(I, X[k]) for k in range(4)])
        np.testing.assert_allclose(F1, G[f'fn1_{fid}'], rtol=2e-13)


def test_protein_energy():
    G = load('protein.npz')
    for pid in G['ids']:
        I = op.ProteinInstance(str(pid), G[f'coor_init_{pid}'], G[f'q_{pid}'], G[f'e_{pid}'], G[f'r_{pid}'],
                               G[f'basis_{pid}'], G[f'eigval_{pid}'])
        E = op.evaluate(I, G[f'x_{pid}'].astype(np.float64))
        np.testing.assert_allclose(E, G[f'energy_{pid}'], rtol=1e-12)


# ----------------------------------------------------------------------------------------- RLEPSO
def _problem_for(fid, dim):
    if fid.isdigit():
        f = int(fid)
        suit = 'bbob' if f < 100 else 'bbob-noisy'
        tr, te = ob.make_suite(suit, dim, 5.0, 'easy')
        I = [i for i in tr + te if i.fid == f][0]
        ev = lambda x, draws: ob.apply_noise(I, ob.func_true(I, x), draws=draws)
        return I, ev, I.optimum, -5.0, 5.0, ob.noise_draw_count(I)
    G = load('protein.npz')
    I = op.ProteinInstance(fid, G[f'coor_init_{fid}'], G[f'q_{fid}'], G[f'e_{fid}'], G[f'r_{fid}'],
                           G[f'basis_{fid}'], G[f'eigval_{fid}'])
    return I, (lambda x, draws: op.energy(I, x)), None, op.LB, op.UB, (0, 0)


def _state_from(G, pre):
    return orl.State(x=G[pre + 'x'], v=G[pre + 'v'], pbest_x=G[pre + 'pbest_x'], c_cost=G[pre + 'c_cost'],
                     pbest=G[pre + 'pbest'], pni=G[pre + 'pni'].astype(np.int32), gbest_x=G[pre + 'gbest_x'],
                     gbest_val=np.float32(G[pre + 'gbest_val']), gbest_idx=int(G[pre + 'gbest_idx']), no_improve=0,
                     fes=int(G[pre + 'fes']), log_index=int(G[pre + 'log_index']),
                     curve=[np.float32(c) for c in G[pre + 'curve']])


def _cmp_state(S, G, pre, tag):
    for k in ('x', 'v', 'pbest_x', 'gbest_x'):
        np.testing.assert_array_equal(getattr(S, k), G[pre + k], err_msg=f'{tag} {k}')
    for k in ('c_cost', 'pbest'):
        np.testing.assert_allclose(getattr(S, k), G[pre + k], rtol=1e-6, atol=1e-10, err_msg=f'{tag} {k}')
    np.testing.assert_array_equal(S.pni, G[pre + 'pni'].astype(np.int32), err_msg=f'{tag} pni')
    assert S.gbest_idx == int(G[pre + 'gbest_idx']), tag
    assert S.fes == int(G[pre + 'fes']), tag
    assert S.log_index == int(G[pre + 'log_index']), tag
    np.testing.assert_allclose(np.array(S.curve, dtype=np.float64), G[pre + 'curve'], rtol=1e-6, atol=1e-10)


def test_rlepso_steps_match_reference():
    G = load('rlepso.npz')
    n_reinit = 0
    for key in G['cases']:
        fid, dim = str(G[f'{key}/fid']), int(G[f'{key}/dim'])
        I, ev, optimum, lb, ub, nc = _problem_for(fid, dim)
        cfg = orl.Config(dim=dim, maxFEs=int(G[f'{key}/maxFEs']), n_logpoint=int(G[f'{key}/n_logpoint']))
        np.random.seed(int(G[f'{key}/seed0']))
        S = orl.init_population(cfg, ev, optimum, lb, ub, orl.NumpyTape(cfg.NP, dim, nc))
        _cmp_state(S, G, f'{key}/s0/', f'{key} init')
        for t in range(1, int(G[f'{key}/n_steps']) + 1):
            S = _state_from(G, f'{key}/s{t - 1}/')          # teacher forcing from the reference's state
            S.pni = G[f'{key}/s{t}/pni_in']
            np.random.seed(int(G[f'{key}/s{t}/seed']))
            aux = {}
            N, feat, reward, done = orl.update(cfg, S, G[f'{key}/s{t}/action'], ev, optimum, lb, ub,
                                               orl.NumpyTape(cfg.NP, dim, nc), aux)
            n_reinit += int(aux['reinit_mask'].any())
            _cmp_state(N, G, f'{key}/s{t}/', f'{key} step {t}')
            assert reward == int(G[f'{key}/s{t}/reward'])
            assert done == bool(G[f'{key}/s{t}/done'])
            np.testing.assert_array_equal(feat, G[f'{key}/s{t}/feature'])
    assert n_reinit >= 5        # the re-initialisation branch is covered


# ----------------------------------------------------------------------------------------- LDE
def _lde_state_from(G, pre):
    from oracle import lde as ol
    ph = G[pre + 'past_histo']
    return ol.State(pop=G[pre + 'pop'], fit=G[pre + 'fit'], hist_sum=ph.sum(0), hist_cnt=ph.shape[0],
                    gbest=np.float32(G[pre + 'gbest']), fes=int(G[pre + 'fes']), log_index=int(G[pre + 'log_index']),
                    curve=[np.float32(c) for c in G[pre + 'curve']])


def _lde_cmp(S, feat, G, pre, tag):
    np.testing.assert_array_equal(S.pop, G[pre + 'pop'], err_msg=f'{tag} pop')
    np.testing.assert_allclose(S.fit, G[pre + 'fit'], rtol=1e-6, atol=1e-10, err_msg=f'{tag} fit')
    ph = G[pre + 'past_histo']
    np.testing.assert_array_equal(S.hist_sum, ph.sum(0), err_msg=f'{tag} histogram sum')
    assert S.hist_cnt == ph.shape[0] and S.fes == int(G[pre + 'fes']) and S.log_index == int(G[pre + 'log_index']), tag
    np.testing.assert_allclose(np.array(S.curve, dtype=np.float64), G[pre + 'curve'], rtol=1e-6, atol=1e-10)
    ref = G[pre + 'feature']
    NP = S.fit.shape[0]
    np.testing.assert_allclose(feat[:NP], ref[:NP], rtol=1e-5, atol=1e-9, err_msg=f'{tag} normalised fitness')
    np.testing.assert_array_equal(feat[NP:NP + 5], ref[NP:NP + 5], err_msg=f'{tag} histogram')
    np.testing.assert_allclose(feat[NP + 5:], ref[NP + 5:], rtol=1e-14, err_msg=f'{tag} mean past histogram')


def test_lde_steps_match_reference():
    """The real LDE_Optimizer (teacher-forced on float32-representable state, seeded np.random + torch) vs the
    oracle replaying the same draws: populations bit-exact, survivor selection / sort order / histograms exact."""
    import torch
    from oracle import lde as ol
    G = load('lde.npz')
    n_done = 0
    for key in G['cases']:
        fid, dim = str(G[f'{key}/fid']), int(G[f'{key}/dim'])
        I, ev, optimum, lb, ub, nc = _problem_for(fid, dim)
        cfg = ol.Config(dim=dim, maxFEs=int(G[f'{key}/maxFEs']), n_logpoint=int(G[f'{key}/n_logpoint']))
        s0 = int(G[f'{key}/seed0'])
        np.random.seed(s0); torch.manual_seed(s0)
        S, feat = ol.init_population(cfg, ev, optimum, lb, ub, ol.NumpyTape(cfg.NP, dim, nc))
        _lde_cmp(S, feat, G, f'{key}/s0/', f'{key} init')
        for t in range(1, int(G[f'{key}/n_steps']) + 1):
            S = _lde_state_from(G, f'{key}/s{t - 1}/')
            s = int(G[f'{key}/s{t}/seed'])
            np.random.seed(s); torch.manual_seed(s)
            aux = {}
            N, feat, reward, done = ol.update(cfg, S, G[f'{key}/s{t}/action'], ev, optimum, lb, ub,
                                              ol.NumpyTape(cfg.NP, dim, nc), aux)
            _lde_cmp(N, feat, G, f'{key}/s{t}/', f'{key} step {t}')
            np.testing.assert_allclose(reward, float(G[f'{key}/s{t}/reward']), rtol=1e-5, atol=1e-12)
            assert done == bool(G[f'{key}/s{t}/done'])
            n_done += int(done)
    assert n_done >= 1


# ----------------------------------------------------------------------------------------- DE-DDQN
def _dd_cmp(S, feat, G, key, t, tag):
    ints, flts = G[f'{key}/ints'][t], G[f'{key}/flts'][t]
    assert [S.pointer, S.gen, S.stagcount, S.fes, S.log_index] == list(ints), f'{tag}: pointer/gen/stagcount/fes/log_index'
    np.testing.assert_allclose([S.c_gbest, S.c_gworst, S.c_prebest], flts, rtol=1e-6, atol=1e-10, err_msg=tag)
    np.testing.assert_array_equal(S.r, G[f'{key}/r'][t], err_msg=f'{tag} r')
    ref = G[f'{key}/feature'][t]
    np.testing.assert_allclose(feat, ref, rtol=2e-5, atol=1e-9, err_msg=f'{tag} features')


def test_deddqn_steps_match_reference():
    """The real DE_DDQN_Optimizer run continuously (state rounded to float32 after each step) vs the oracle on the
    same seeds and actions: replaced rows / gbest / prebest bit-exact, all 99 features, credit history, window."""
    from oracle import deddqn as od
    G = load('deddqn.npz')
    n_total = n_gen2 = n_evict = n_done = 0
    for key in G['cases']:
        fid, dim = str(G[f'{key}/fid']), int(G[f'{key}/dim'])
        I, ev, optimum, lb, ub, nc = _problem_for(fid, dim)
        cfg = od.Config(dim=dim, maxFEs=int(G[f'{key}/maxFEs']), n_logpoint=int(G[f'{key}/n_logpoint']))
        s0 = int(G[f'{key}/seed0'])
        np.random.seed(s0)
        S, feat = od.init_population(cfg, ev, optimum, lb, ub, od.NumpyTape(cfg.NP, dim, nc))
        np.testing.assert_array_equal(S.X, G[f'{key}/X0'])
        np.testing.assert_allclose(S.cost, G[f'{key}/cost0'], rtol=1e-6, atol=1e-10)
        _dd_cmp(S, feat, G, key, 0, f'{key} init')
        snaps = set(int(t) for t in G[f'{key}/snap_steps'])
        for t in range(1, int(G[f'{key}/n_steps']) + 1):
            p = S.pointer
            np.random.seed(s0 + t)
            aux = {}
            S, feat, reward, done = od.update(cfg, S, int(G[f'{key}/action'][t - 1]), ev, optimum, lb, ub,
                                              od.NumpyTape(cfg.NP, dim, nc), aux)
            tag = f'{key} step {t}'
            _dd_cmp(S, feat, G, key, t, tag)
            np.testing.assert_array_equal(S.X[p], G[f'{key}/row'][t - 1], err_msg=f'{tag} row')
            np.testing.assert_allclose(S.cost[p], G[f'{key}/row_cost'][t - 1], rtol=1e-6, atol=1e-10)
            np.testing.assert_array_equal(S.X_gbest, G[f'{key}/X_gbest'][t - 1], err_msg=f'{tag} X_gbest')
            np.testing.assert_array_equal(S.prebest(), G[f'{key}/X_prebest'][t - 1], err_msg=f'{tag} X_prebest')
            np.testing.assert_allclose(reward, G[f'{key}/reward'][t - 1], rtol=1e-5, atol=1e-12)
            assert done == bool(G[f'{key}/done'][t - 1]), tag
            if t in snaps:
                np.testing.assert_array_equal(S.X, G[f'{key}/snapX_{t}'], err_msg=f'{tag} X snapshot')
                np.testing.assert_allclose(S.cost, G[f'{key}/snapC_{t}'], rtol=1e-6, atol=1e-10)
            n_total += 1; n_gen2 += int(S.gen >= 3); n_evict += int(S.win_len >= od.W_); n_done += int(done)
        np.testing.assert_allclose(np.array(S.curve, dtype=np.float64), G[f'{key}/curve'], rtol=1e-6, atol=1e-10)
    assert n_total >= 900 and n_gen2 >= 100 and n_evict >= 300 and n_done >= 1


# ----------------------------------------------------------------------------------------- Random_search
def test_random_search_episodes_match_reference():
    """`Random_search.run_episode` of the reference (float32 storage contract, np.random.seed(5)) vs the oracle replaying the
    same draws, incl. a noisy function (draw order: positions, then the noise model's draws inside problem.eval)."""
    from oracle import random_search as ors
    G = load('rs.npz')
    for fid, dim in ((1, 10), (15, 10), (8, 30), (102, 10)):
        I, ev, optimum, lb, ub, nc = _problem_for(str(fid), dim)
        np.random.seed(5)
        r = ors.run_episode(dim, 2000 * dim, 50, 2000 * dim // 50, ev, optimum, lb, ub, ors.NumpyTape(100, dim, nc))
        assert r['fes'] == int(G[f'f32/{fid}_{dim}/fes'])
        np.testing.assert_allclose(np.array(r['cost'], dtype=np.float64), G[f'f32/{fid}_{dim}/cost'], rtol=1e-6, atol=1e-10)


# ----------------------------------------------------------------------------------------- GLEET
def _gleet_state_from(G, pre):
    from oracle import gleet as og
    return og.State(x=G[pre + 'x'], v=G[pre + 'v'], pbest_x=G[pre + 'pbest_x'], c_cost=G[pre + 'c_cost'], pbest=G[pre + 'pbest'],
                    pni=G[pre + 'pni'].astype(np.int32), gbest_x=G[pre + 'gbest_x'], gbest_val=np.float32(G[pre + 'gbest_val']),
                    gbest_idx=int(G[pre + 'gbest_idx']), no_improve=int(G[pre + 'no_improve']), w=float(G[pre + 'w']),
                    max_cost=np.float32(G[pre + 'max_cost']), pbest_feature=G[pre + 'pbest_feature'],
                    gbest_feature=G[pre + 'gbest_feature'], fes=int(G[pre + 'fes']), log_index=int(G[pre + 'log_index']),
                    curve=[np.float32(c) for c in G[pre + 'curve']])


def _gleet_cmp(S, state, G, pre, tag):
    for k in ('x', 'v', 'pbest_x', 'gbest_x'):
        np.testing.assert_array_equal(getattr(S, k), G[pre + k], err_msg=f'{tag} {k}')
    for k in ('c_cost', 'pbest'):
        np.testing.assert_allclose(getattr(S, k), G[pre + k], rtol=1e-6, atol=1e-10, err_msg=f'{tag} {k}')
    np.testing.assert_array_equal(S.pni, G[pre + 'pni'], err_msg=f'{tag} pni')
    assert S.gbest_idx == int(G[pre + 'gbest_idx']) and S.no_improve == int(G[pre + 'no_improve']), tag
    assert S.fes == int(G[pre + 'fes']) and S.log_index == int(G[pre + 'log_index']), tag
    assert S.w == float(G[pre + 'w']), tag                                   # the same float64 subtraction chain
    # the reference computes its features from the float64 positions BEFORE they are rounded to the float32 store, the
    # engine from the stored float32 values: cost / counter / distance features agree to float32 rounding, the cosine
    # feature (col 8: a dot product with cancellation, range [-1, 1]) to 5e-6 absolute
    atol = np.where(np.arange(27) % 9 == 8, 5e-6, 1e-9)
    close = lambda a, b, at: np.all(np.abs(np.asarray(a, dtype=np.float64) - b) <= at + 2e-6 * np.abs(b))
    assert close(S.pbest_feature, G[pre + 'pbest_feature'], atol[None, :9]), f'{tag} pbest_feature'
    assert close(S.gbest_feature, G[pre + 'gbest_feature'], atol[:9]), f'{tag} gbest_feature'
    assert np.all(np.abs(state - G[pre + 'state']) <= atol[None, :] + 2e-6 * np.abs(G[pre + 'state'])), f'{tag} state [NP,27]'
    np.testing.assert_allclose(np.array(S.curve, dtype=np.float64), G[pre + 'curve'], rtol=1e-6, atol=1e-10)


def test_gleet_steps_match_reference():
    """The real GLEET_Optimizer (teacher-forced on float32-representable state, seeded np.random) vs the oracle replaying
    the same draws: positions / velocities / pbest bit-exact, the 27 features per particle within float32 rounding."""
    from oracle import gleet as og
    G = load('gleet.npz')
    for key in G['cases']:
        fid, dim = str(G[f'{key}/fid']), int(G[f'{key}/dim'])
        I, ev, optimum, lb, ub, nc = _problem_for(fid, dim)
        cfg = og.Config(dim=dim, maxFEs=int(G[f'{key}/maxFEs']), n_logpoint=int(G[f'{key}/n_logpoint']))
        np.random.seed(int(G[f'{key}/seed0']))
        S, state = og.init_population(cfg, ev, optimum, lb, ub, og.NumpyTape(cfg.NP, dim, nc))
        _gleet_cmp(S, state, G, f'{key}/s0/', f'{key} init')
        for t in range(1, int(G[f'{key}/n_steps']) + 1):
            S = _gleet_state_from(G, f'{key}/s{t - 1}/')
            np.random.seed(int(G[f'{key}/s{t}/seed']))
            N, state, reward, done = og.update(cfg, S, G[f'{key}/s{t}/action'], ev, optimum, lb, ub, og.NumpyTape(cfg.NP, dim, nc))
            _gleet_cmp(N, state, G, f'{key}/s{t}/', f'{key} step {t}')
            np.testing.assert_allclose(reward, float(G[f'{key}/s{t}/reward']), rtol=1e-6, atol=1e-9)
            assert done == bool(G[f'{key}/s{t}/done'])


# ----------------------------------------------------------------------------------------- landscape features
def test_landscape_features_match_reference():
    """cal_fdc / cal_rie / cal_acf / cal_nop of the reference's dedqn_optimizer.py and the DFDC / DRIE classifications of
    rl_hpsde_optimizer.py on seeded walks (incl. stepped and plateau costs) vs the oracle."""
    from oracle import landscape as ol
    G = load('landscape.npz')
    n_bool = 0
    for key in G['cases']:
        x, c, ref = G[f'{key}/x'], G[f'{key}/c'], G[f'{key}/ref']
        f = ol.features(x, c, ref)
        np.testing.assert_allclose(f[:4], G[f'{key}/dedqn'], rtol=1e-9, atol=1e-12, err_msg=key)
        a, b = G[f'{key}/hpsde_bool']
        if a >= 0 and abs(f[4] - 0.15) > 1e-9:
            assert int(0.15 < f[4] <= 1) == a, key; n_bool += 1            # rl_hpsde_optimizer.py:175-180
        if b >= 0 and abs(f[5] - 0.5) > 1e-9:
            assert int(0.5 <= f[5] <= 1) == b, key; n_bool += 1             # :212-217
    assert n_bool >= 40


# ----------------------------------------------------------------------------------------- RL_PSO
def _rlpso_s0(G, key):
    from oracle import rl_pso as orp
    pre = f'{key}/s0/'
    return orp.State(x=G[pre + 'x'], v=G[pre + 'v'], pbest_x=G[pre + 'pbest_x'], c_cost=G[pre + 'c_cost'], pbest=G[pre + 'pbest'],
                     gbest_x=G[pre + 'gbest_x'], gbest_val=np.float32(G[pre + 'gbest_val']), gbest_idx=int(G[pre + 'gbest_idx']), w=0.9,
                     max_cost=np.float32(G[pre + 'max_cost']), cur_index=0, fes=len(G[pre + 'c_cost']), log_index=1,
                     curve=[np.float32(G[pre + 'gbest_val'])])


def _rlpso_cmp_step(N, state, reward, done, G, key, t, tag):
    j = (t - 1) % len(N.c_cost)
    k = t - 1
    np.testing.assert_array_equal(N.x[j], G[f'{key}/row_x'][k], err_msg=f'{tag} x row')
    np.testing.assert_array_equal(N.v[j], G[f'{key}/row_v'][k], err_msg=f'{tag} v row')
    np.testing.assert_array_equal(N.pbest_x[j], G[f'{key}/row_pbest_x'][k], err_msg=f'{tag} pbest_x row')
    np.testing.assert_array_equal(N.gbest_x, G[f'{key}/gbest_x'][k], err_msg=f'{tag} gbest_x')
    np.testing.assert_allclose(N.c_cost[j], G[f'{key}/row_c_cost'][k], rtol=1e-6, atol=1e-10, err_msg=f'{tag} cost')
    np.testing.assert_allclose(N.pbest[j], G[f'{key}/row_pbest'][k], rtol=1e-6, atol=1e-10)
    np.testing.assert_allclose(float(N.gbest_val), G[f'{key}/gbest_val'][k], rtol=1e-6, atol=1e-10)
    assert N.gbest_idx == int(G[f'{key}/gbest_idx'][k]) and N.w == float(G[f'{key}/w'][k]), tag
    assert N.fes == int(G[f'{key}/fes'][k]) and N.log_index == int(G[f'{key}/log_index'][k]) and len(N.curve) == int(G[f'{key}/curve_len'][k]), tag
    np.testing.assert_allclose(float(N.curve[-1]), G[f'{key}/curve_last'][k], rtol=1e-6, atol=1e-10)
    np.testing.assert_allclose(state, G[f'{key}/state'][k], rtol=1e-6, atol=1e-7, err_msg=f'{tag} state')
    np.testing.assert_allclose(reward, G[f'{key}/reward'][k], rtol=2e-5, atol=1e-9, err_msg=f'{tag} reward')
    assert done == bool(G[f'{key}/done'][k]), tag


def test_rl_pso_steps_match_reference():
    """230 consecutive single-particle steps of the real RL_PSO_Optimizer per case (float32 storage contract, seeded
    np.random) vs the oracle replaying the same draws: the updated row, gbest and the inertia chain bit-exact."""
    from oracle import rl_pso as orp
    G = load('rl_pso.npz')
    for key in G['cases']:
        fid, dim = str(G[f'{key}/fid']), int(G[f'{key}/dim'])
        I, ev, optimum, lb, ub, nc = _problem_for(fid, dim)
        cfg = orp.Config(dim=dim, maxFEs=int(G[f'{key}/maxFEs']), n_logpoint=int(G[f'{key}/n_logpoint']))
        np.random.seed(int(G[f'{key}/seed0']))
        S, state = orp.init_population(cfg, ev, optimum, lb, ub, orp.NumpyTape(cfg.NP, dim, nc))
        S0 = _rlpso_s0(G, key)
        for k in ('x', 'v', 'pbest_x', 'gbest_x'):
            np.testing.assert_array_equal(getattr(S, k), getattr(S0, k), err_msg=f'{key} init {k}')
        np.testing.assert_allclose(S.c_cost, S0.c_cost, rtol=1e-6, atol=1e-10)
        np.testing.assert_allclose(float(S.max_cost), float(S0.max_cost), rtol=1e-6)
        np.testing.assert_allclose(state, G[f'{key}/s0/state'], rtol=1e-6, atol=1e-7)
        S = S0                                                       # carry the reference's (float32) costs forward
        for t in range(1, int(G['n_steps']) + 1):
            np.random.seed(int(G[f'{key}/seed'][t - 1]))
            S, state, reward, done = orp.update(cfg, S, G[f'{key}/action'][t - 1], ev, optimum, lb, ub, orp.NumpyTape(cfg.NP, dim, nc))
            _rlpso_cmp_step(S, state, reward, done, G, key, t, f'{key} step {t}')
            # teacher forcing of the float32 costs (the reference's value may differ in the last float32 bit)
            j = (t - 1) % cfg.NP
            S.c_cost[j] = G[f'{key}/row_c_cost'][t - 1]; S.pbest[j] = G[f'{key}/row_pbest'][t - 1]
            S.gbest_val = np.float32(G[f'{key}/gbest_val'][t - 1])


# ------------------------------------------------------------------------------------------ QLPSO
def _qlpso_s0(G, key):
    from oracle import qlpso as oq
    pre = f'{key}/s0/'
    return oq.State(x=G[pre + 'x'], v=G[pre + 'v'], pbest_x=G[pre + 'pbest_x'], cost=G[pre + 'cost'],
                    gbest_cost=np.float32(G[pre + 'gbest_cost']), diversity=float(G[pre + 'diversity']),
                    q_state=G[pre + 'q_state'].astype(np.int32), pointer=0, fes=len(G[pre + 'cost']), log_index=1,
                    curve=[np.float32(G[pre + 'gbest_cost'])])


def test_qlpso_steps_match_reference():
    """100 consecutive single-particle steps of the real QLPSO_Optimizer per case (float32 storage contract, seeded
    np.random) vs the oracle replaying the same draws: the updated row and the float64 diversity bit-exact, the integer
    reward and the returned Q-state equal, costs to 1e-6."""
    from oracle import qlpso as oq
    G = load('qlpso.npz')
    seen_rewards = set()
    for key in G['cases']:
        fid, dim = str(G[f'{key}/fid']), int(G[f'{key}/dim'])
        I, ev, optimum, lb, ub, nc = _problem_for(fid, dim)
        cfg = oq.Config(dim=dim, maxFEs=int(G[f'{key}/maxFEs']), n_logpoint=int(G[f'{key}/n_logpoint']))
        np.random.seed(int(G[f'{key}/seed0']))
        S, qs = oq.init_population(cfg, ev, optimum, lb, ub, oq.NumpyTape(cfg.NP, dim, nc))
        S0 = _qlpso_s0(G, key)
        for k in ('x', 'v', 'pbest_x', 'q_state'):
            np.testing.assert_array_equal(getattr(S, k), getattr(S0, k), err_msg=f'{key} init {k}')
        np.testing.assert_allclose(S.cost, S0.cost, rtol=1e-6, atol=1e-10)
        assert S.diversity == S0.diversity and qs == int(G[f'{key}/s0/ret_state'])
        S = S0                                                       # carry the reference's (float32) costs forward
        for t in range(1, int(G['n_steps']) + 1):
            k = t - 1
            j = k % cfg.NP
            tag = f'{key} step {t}'
            np.random.seed(int(G[f'{key}/seed'][k]))
            S, qs, reward, done = oq.update(cfg, S, G[f'{key}/action'][k], ev, optimum, lb, ub, oq.NumpyTape(cfg.NP, dim, nc))
            np.testing.assert_array_equal(S.x[j], G[f'{key}/row_x'][k], err_msg=f'{tag} x row')
            np.testing.assert_array_equal(S.v[j], G[f'{key}/row_v'][k], err_msg=f'{tag} v row')
            np.testing.assert_array_equal(S.pbest_x[j], G[f'{key}/row_pbest_x'][k], err_msg=f'{tag} pbest row')
            np.testing.assert_allclose(S.cost[j], G[f'{key}/row_cost'][k], rtol=1e-6, atol=1e-10, err_msg=f'{tag} cost')
            np.testing.assert_allclose(float(S.gbest_cost), G[f'{key}/gbest_cost'][k], rtol=1e-6, atol=1e-10)
            assert S.diversity == float(G[f'{key}/diversity'][k]), tag
            assert reward == int(G[f'{key}/reward'][k]) and qs == int(G[f'{key}/next_state'][k]) and done == bool(G[f'{key}/done'][k]), tag
            assert S.fes == int(G[f'{key}/fes'][k]) and S.log_index == int(G[f'{key}/log_index'][k]) and len(S.curve) == int(G[f'{key}/curve_len'][k]), tag
            np.testing.assert_allclose(float(S.curve[-1]), G[f'{key}/curve_last'][k], rtol=1e-6, atol=1e-10)
            seen_rewards.add(reward)
            S.cost[j] = G[f'{key}/row_cost'][k]; S.gbest_cost = np.float32(G[f'{key}/gbest_cost'][k])
    assert seen_rewards == {2, 1, 0, -2}


def test_gleet_and_rl_pso_policy_restatements_match_reference():
    """oracle.agents' head-by-head GLEET actor and the RL_PSO policy network against forwards of the unmodified reference
    (agents2.npz), and a bounded oracle GLEET rollout driven by the attention actor (bench.py's CPU arm for --workload gleet)."""
    import torch
    from oracle import agents as oag, rollout as orr
    G = np.load(os.path.join(GOLDEN, 'agents2.npz'))
    for grp in ('gleet0', 'gleet'):
        w = oag.load_weights(G, grp)
        for name in ('a', 'b', 'r'):
            mu, sigma, logits = oag.gleet_distribution(w, G[f'{grp}_kat/{name}/x'])
            np.testing.assert_allclose(logits.numpy(), G[f'{grp}_kat/{name}/logits'], rtol=2e-4, atol=2e-5)
            np.testing.assert_allclose(mu.numpy(), G[f'{grp}_kat/{name}/mu'][..., 0], rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(sigma.numpy(), G[f'{grp}_kat/{name}/sigma'][..., 0], rtol=1e-4, atol=1e-5)
    a = oag.gleet_act(w, G['gleet_kat/a/x'][0])
    assert a.shape == (100,) and a.dtype == np.float32 and a.min() >= 0 and a.max() <= 1
    w2 = oag.load_weights(G, 'rl_pso')
    mu, sigma = oag.rl_pso_distribution(w2, G['rl_pso_kat/x'])
    np.testing.assert_allclose(mu.numpy(), G['rl_pso_kat/mu'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(sigma.numpy(), G['rl_pso_kat/sigma'], rtol=1e-5, atol=1e-6)
    assert oag.rl_pso_act(w2, G['rl_pso_kat/x']).shape == (9, 1)
    tr, te = ob.make_suite('bbob', 10, 5.0)
    r = orr.gleet_rollout((tr + te)[0], 10, 20000, w, run_seed=3, max_gens=4)
    assert r['gens'] == 4 and r['fes'] == 500 and np.isfinite(r['cost']).all()
    r2 = orr.gleet_rollout((tr + te)[0], 10, 20000, w, run_seed=3, max_gens=4)
    assert r2['cost'] == r['cost']                                           # seeded -> repeatable


# ------------------------------------------------------------------------------------------ DEDQN
def test_dedqn_steps_match_reference():
    """130 consecutive single-individual steps of the real DEDQN_Optimizer per case (float32 storage contract, np.random re-seeded before
    every step) vs the oracle carrying its OWN state and replaying the same draws: replaced rows / gbest bit-exact, the 4 landscape
    features, the survival-based reward, fes accounting (200 charged evaluations per step) and the cost-curve protocol past maxFEs."""
    from oracle import dedqn as od
    G = load('dedqn.npz')
    for key in G['cases']:
        fid, dim = str(G[f'{key}/fid']), int(G[f'{key}/dim'])
        I, ev, optimum, lb, ub, nc = _problem_for(fid, dim)
        cfg = od.Config(dim=dim, maxFEs=int(G[f'{key}/maxFEs']), n_logpoint=int(G[f'{key}/n_logpoint']))
        np.random.seed(int(G[f'{key}/seed0']))
        S, state = od.init_population(cfg, ev, optimum, lb, ub, od.NumpyTape(cfg.NP, dim, nc))
        np.testing.assert_array_equal(S.pop, G[f'{key}/s0/pop'], err_msg=f'{key} init pop')
        np.testing.assert_allclose(S.cost, G[f'{key}/s0/cost'], rtol=1e-6, atol=1e-10)
        np.testing.assert_array_equal(S.best(), G[f'{key}/s0/gbest'])
        np.testing.assert_allclose(state, G[f'{key}/s0/state'], rtol=1e-6, atol=1e-9)
        assert S.fes == int(G[f'{key}/s0/fes']) == 2 * cfg.NP
        S.cost = G[f'{key}/s0/cost'].copy(); S.gbest_cost = np.float32(G[f'{key}/s0/gbest_cost'])       # carry the reference's float32 costs
        seen_done = False
        for t in range(1, int(G['n_steps']) + 1):
            np.random.seed(int(G[f'{key}/seed'][t - 1]))
            j = S.pointer
            S, state, reward, done = od.update(cfg, S, int(G[f'{key}/action'][t - 1]), ev, optimum, lb, ub, od.NumpyTape(cfg.NP, dim, nc))
            tag = f'{key} step {t}'
            np.testing.assert_array_equal(S.pop[j], G[f'{key}/row'][t - 1], err_msg=tag)
            np.testing.assert_allclose(S.cost[j], G[f'{key}/row_cost'][t - 1], rtol=1e-6, atol=1e-10, err_msg=tag)
            assert S.survival[j] == G[f'{key}/row_survival'][t - 1], tag
            np.testing.assert_array_equal(S.best(), G[f'{key}/gbest'][t - 1], err_msg=tag)
            np.testing.assert_allclose(float(S.gbest_cost), float(G[f'{key}/gbest_cost'][t - 1]), rtol=1e-6, atol=1e-10, err_msg=tag)
            np.testing.assert_allclose(state, G[f'{key}/state'][t - 1], rtol=2e-5, atol=1e-8, err_msg=tag)
            np.testing.assert_allclose(reward, G[f'{key}/reward'][t - 1], rtol=1e-12, err_msg=tag)
            assert done == bool(G[f'{key}/done'][t - 1]) and S.fes == int(G[f'{key}/fes'][t - 1]) == 200 + 200 * t, tag
            assert S.log_index == int(G[f'{key}/log_index'][t - 1]) and len(S.curve) == int(G[f'{key}/curve_len'][t - 1]), tag
            np.testing.assert_allclose(float(S.curve[-1]), float(G[f'{key}/curve_last'][t - 1]), rtol=1e-6, atol=1e-10, err_msg=tag)
            S.cost[j] = G[f'{key}/row_cost'][t - 1]; S.gbest_cost = np.float32(G[f'{key}/gbest_cost'][t - 1])   # last-bit teacher forcing of costs
            seen_done |= done
        np.testing.assert_array_equal(S.pop, G[f'{key}/final_pop'], err_msg=f'{key} final pop')
        np.testing.assert_array_equal(S.survival, G[f'{key}/final_survival'])
        assert seen_done == (key == 'f1_d10')
    with pytest.raises(ValueError):
        od.update(cfg, S, 3, ev, optimum, lb, ub, od.NumpyTape(cfg.NP, dim, nc))


# ------------------------------------------------------------------------------------------ RL_HPSDE
def test_rl_hpsde_generations_match_reference():
    """Consecutive generations of the real RL_HPSDE_Optimizer per case (float32 storage contract, np.random re-seeded before every
    generation) vs the oracle carrying its OWN state and replaying the same draws (incl. scipy's cauchy / levy samplers): sorted
    population and costs, success-history memories, population-size schedule, landscape state, reward, accounting, curve protocol; one
    case runs to the end of its budget."""
    from oracle import rl_hpsde as oh
    G = load('rl_hpsde.npz')
    seen_states, seen_done, n_tie_swaps = set(), False, 0
    for key in G['cases']:
        fid, dim = str(G[f'{key}/fid']), int(G[f'{key}/dim'])
        I, ev, optimum, lb, ub, nc = _problem_for(fid, dim)
        cfg = oh.Config(dim=dim, maxFEs=int(G[f'{key}/maxFEs']), n_logpoint=int(G[f'{key}/n_logpoint']))
        np.random.seed(int(G[f'{key}/seed0']))
        S, state = oh.init_population(cfg, ev, optimum, lb, ub, oh.NumpyTape(dim, nc))
        np.testing.assert_array_equal(S.group, G[f'{key}/s0/group'], err_msg=f'{key} init group')
        np.testing.assert_allclose(S.cost, G[f'{key}/s0/cost'], rtol=1e-6, atol=1e-10)
        assert state == int(G[f'{key}/s0/state']) and S.fes == int(G[f'{key}/s0/fes']) == 18 * dim + 201
        S.cost = G[f'{key}/s0/cost'].copy()
        for t in range(1, int(G[f'{key}/n_steps']) + 1):
            np.random.seed(int(G[f'{key}/seed'][t - 1]))
            S, state, reward, done = oh.update(cfg, S, int(G[f'{key}/action'][t - 1]), ev, optimum, lb, ub, oh.NumpyTape(dim, nc))
            tag = f'{key} generation {t}'
            assert S.NP == int(G[f'{key}/NP'][t - 1]), tag
            np.testing.assert_allclose(S.cost, G[f'{key}/cost/{t}'], rtol=1e-6, atol=1e-10, err_msg=tag)
            ref_group = G[f'{key}/group/{t}']
            if not np.array_equal(S.group, ref_group):                     # np.argsort (quicksort) vs the oracle's stable order of TIED costs
                for lo in np.flatnonzero(np.r_[True, S.cost[1:] != S.cost[:-1]]):
                    hi = lo + 1
                    while hi < S.NP and S.cost[hi] == S.cost[lo]:
                        hi += 1
                    rows = lambda A: sorted(map(tuple, A[lo:hi].tolist()))
                    assert rows(S.group) == rows(ref_group), tag
                n_tie_swaps += 1
                S.group = ref_group.copy()
            np.testing.assert_allclose(S.MF, G[f'{key}/MF'][t - 1], rtol=1e-5, atol=1e-9, err_msg=tag)
            np.testing.assert_allclose(S.MCr, G[f'{key}/MCr'][t - 1], rtol=1e-5, atol=1e-9, err_msg=tag)
            assert S.k == int(G[f'{key}/k'][t - 1]) and state == int(G[f'{key}/state'][t - 1]), tag
            np.testing.assert_allclose(reward, G[f'{key}/reward'][t - 1], rtol=1e-12, err_msg=tag)
            assert done == bool(G[f'{key}/done'][t - 1]) and S.fes == int(G[f'{key}/fes'][t - 1]), tag
            assert S.log_index == int(G[f'{key}/log_index'][t - 1]) and len(S.curve) == int(G[f'{key}/curve_len'][t - 1]), tag
            np.testing.assert_allclose(float(S.curve[-1]), float(G[f'{key}/curve_last'][t - 1]), rtol=1e-6, atol=1e-10, err_msg=tag)
            S.cost = G[f'{key}/cost/{t}'].copy(); S.gbest = np.float32(G[f'{key}/gbest'][t - 1])       # last-bit teacher forcing of the costs
            S.MF, S.MCr = G[f'{key}/MF'][t - 1].copy(), G[f'{key}/MCr'][t - 1].copy()
            seen_states.add(state); seen_done |= done
    assert seen_done and seen_states == {0, 1, 2, 3} and n_tie_swaps >= 1          # tied costs do occur (plateaus, converged swarms)
    with pytest.raises(ValueError):
        oh.update(cfg, S, 4, ev, optimum, lb, ub, oh.NumpyTape(dim, nc))


# ------------------------------------------------------------------------------------------ L2L
def test_l2l_episodes_match_reference():
    """Whole 100-step episodes of the real L2L agent + optimizer (shipped weights, dim 10; no random draws for noise-free functions) vs the
    oracle's LSTM-with-projection restatement and update rule: every cost the episode saw, the returned curve, fes and return."""
    from oracle import l2l as ol
    G = load('l2l.npz')
    import torch
    w = {k[4:]: torch.as_tensor(G[k]) for k in G.files if k.startswith('net/')}
    for fid in G['fids']:
        if fid > 100:
            continue                                                   # noisy: the reference's noise comes from np.random inside eval
        I, ev, optimum, lb, ub, nc = _problem_for(str(fid), 10)
        r = ol.rollout(w, 10, ev, optimum, lb, ub)
        np.testing.assert_allclose(r['y'], G[f'f{fid}/y'], rtol=2e-6, err_msg=f'f{fid}')
        np.testing.assert_allclose(r['cost'], G[f'f{fid}/cost'], rtol=2e-6)
        assert r['fes'] == int(G[f'f{fid}/fes']) == 100 and len(r['cost']) == 50
        np.testing.assert_allclose(r['ret'], float(G[f'f{fid}/ret']), rtol=1e-6)


# ------------------------------------------------------------------------------------------ NRLPSO (oracle only: groundwork for a CUDA path)
def test_nrlpso_steps_match_reference():
    """260 consecutive single-particle steps of the real NRLPSO_Optimizer per case (float32 storage contract, np.random re-seeded before
    every step) vs the oracle carrying its OWN state: the moved row, its velocity, personal / global best incl. the gbest view aliasing and
    the never-refreshed pbest costs, neighbourhood mutants, rewards, variable evaluation counts, the next particle's state."""
    from oracle import nrlpso as on
    G = load('nrlpso.npz')
    n_mut = 0
    for key in G['cases']:
        fid, dim = str(G[f'{key}/fid']), int(G[f'{key}/dim'])
        I, ev, optimum, lb, ub, nc = _problem_for(fid, dim)
        cfg = on.Config(dim=dim, maxFEs=int(G[f'{key}/maxFEs']), n_logpoint=int(G[f'{key}/n_logpoint']))
        np.random.seed(int(G[f'{key}/seed0']))
        S, state = on.init_population(cfg, ev, optimum, lb, ub, on.NumpyTape(dim, nc))
        np.testing.assert_array_equal(S.pop, G[f'{key}/s0/pop'], err_msg=f'{key} init')
        np.testing.assert_allclose(S.cost, G[f'{key}/s0/cost'], rtol=1e-6, atol=1e-10)
        assert state == int(G[f'{key}/s0/state']) and S.r_w == float(G[f'{key}/s0/r_w'])
        S.cost = G[f'{key}/s0/cost'].copy(); S.pbest_cost = S.cost.copy(); S.gbest_cost = np.float32(S.cost.min())
        for t in range(1, int(G['n_steps']) + 1):
            np.random.seed(int(G[f'{key}/seed'][t - 1]))
            j, fes0 = S.pointer, S.fes
            S, state, reward, done = on.update(cfg, S, int(G[f'{key}/action'][t - 1]), ev, optimum, lb, ub, on.NumpyTape(dim, nc))
            tag = f'{key} step {t}'
            np.testing.assert_array_equal(S.vel[j], G[f'{key}/row_vel'][t - 1], err_msg=tag)
            np.testing.assert_array_equal(S.pop[j], G[f'{key}/row_pop'][t - 1], err_msg=tag)
            np.testing.assert_allclose(S.cost[j], G[f'{key}/row_cost'][t - 1], rtol=1e-6, atol=1e-10, err_msg=tag)
            np.testing.assert_array_equal(S.pbest_pos[j], G[f'{key}/row_pbest_pos'][t - 1], err_msg=tag)
            np.testing.assert_allclose(S.pbest_cost[j], G[f'{key}/row_pbest_cost'][t - 1], rtol=1e-6, atol=1e-10, err_msg=tag)
            np.testing.assert_array_equal(S.gbest(), G[f'{key}/gbest_pos'][t - 1], err_msg=tag)
            np.testing.assert_allclose(float(S.gbest_cost), float(G[f'{key}/gbest_cost'][t - 1]), rtol=1e-6, atol=1e-10, err_msg=tag)
            assert S.stag[j] == int(G[f'{key}/stag'][t - 1]) and state == int(G[f'{key}/state'][t - 1]) and reward == int(G[f'{key}/reward'][t - 1]), tag
            assert done == bool(G[f'{key}/done'][t - 1]) and S.fes == int(G[f'{key}/fes'][t - 1]) and S.log_index == int(G[f'{key}/log_index'][t - 1]), tag
            assert len(S.curve) == int(G[f'{key}/curve_len'][t - 1]) and abs(S.w - float(G[f'{key}/w'][t - 1])) <= 1e-15, tag
            n_mut += (S.fes - fes0 - 1) // 2
            if f'{key}/pop/{t}' in G:
                np.testing.assert_array_equal(S.pop, G[f'{key}/pop/{t}'], err_msg=tag)
                np.testing.assert_allclose(S.cost, G[f'{key}/cost/{t}'], rtol=1e-6, atol=1e-10, err_msg=tag)
                S.cost = G[f'{key}/cost/{t}'].copy()
            S.cost[j] = G[f'{key}/row_cost'][t - 1]; S.pbest_cost[j] = G[f'{key}/row_pbest_cost'][t - 1]      # last-bit teacher forcing of costs
            S.gbest_cost = np.float32(G[f'{key}/gbest_cost'][t - 1])
    assert n_mut >= 40


# ------------------------------------------------------------------------------------------ LES
def test_les_generations_match_reference():
    """40 consecutive generations of the real LES_Optimizer per case (stepped with its own skip_step = 1, np.random re-seeded before every
    generation) vs the oracle carrying its OWN float64 state: population, raw costs, mean / step size, the six evolution paths, gbest,
    evaluation count and log index."""
    from oracle import les as ol
    G = load('les.npz')
    attn, mlp = G['attn'], G['mlp']
    assert len(attn) == ol.N_ATTN and len(mlp) == ol.N_MLP
    for key in G['cases']:
        fid, dim = str(G[f'{key}/fid']), int(G[f'{key}/dim'])
        I, ev, optimum, lb, ub, nc = _problem_for(fid, dim)
        maxFEs, log_interval = int(G[f'{key}/maxFEs']), int(G[f'{key}/log_interval'])
        np.random.seed(int(G[f'{key}/seed0']))
        S = ol.init_population(dim, maxFEs, ev, lb, ub)
        np.testing.assert_array_equal(S.parents, G[f'{key}/s0/parents']); np.testing.assert_array_equal(S.mu, G[f'{key}/s0/mu'])
        np.testing.assert_allclose(S.cost, G[f'{key}/s0/cost'], rtol=1e-12)
        for t in range(1, int(G['n_gen']) + 1):
            np.random.seed(int(G[f'{key}/seed'][t - 1]))
            S = ol.generation(S, attn, mlp, dim, maxFEs, log_interval, ev, optimum, lb, ub)
            tag = f'{key} generation {t}'
            for k in ('parents', 'mu', 'sigma', 'Pc', 'Ps'):
                np.testing.assert_allclose(getattr(S, k), G[f'{key}/{k}'][t - 1], rtol=1e-9, atol=1e-12, err_msg=f'{tag} {k}')
            np.testing.assert_allclose(S.cost, G[f'{key}/cost'][t - 1], rtol=1e-9, err_msg=tag)
            np.testing.assert_allclose(S.gbest, float(G[f'{key}/gbest'][t - 1]), rtol=1e-12, err_msg=tag)
            assert S.fes == int(G[f'{key}/fes'][t - 1]) == 16 * (t + 1) and S.t == t and S.log_index == int(G[f'{key}/log_index'][t - 1]), tag


def test_oracle_rollouts_of_the_composed_backbones_run_bounded_samples():
    """CPU-arm rollouts (oracle/rollout.py) of DEDQN / RL_HPSDE / NRLPSO / LES with the shipped or golden policies: bounded samples run,
    charge the reference's evaluation counts and are repeatable under a seed (what `bench.py --impl reference` times for the six native
    workloads is available for the composed ones too)."""
    from oracle import agents as oag, rollout as orr
    A2, GL = load('agents2.npz'), load('les.npz')
    tr, te = ob.make_suite('bbob', 10, 5.0)
    inst = [i for i in tr + te if i.fid == 15][0]
    w_dq = {k[len('dedqn/'):]: A2[k] for k in A2.files if k.startswith('dedqn/')}
    r = orr.dedqn_rollout(inst, 10, 20000, w_dq, run_seed=1, max_gens=5)
    assert r['gens'] == 5 and r['fes'] == 200 + 5 * 200
    r = orr.rl_hpsde_rollout(inst, 10, 20000, A2['rl_hpsde_q'], run_seed=1, max_gens=3)
    r2 = orr.rl_hpsde_rollout(inst, 10, 20000, A2['rl_hpsde_q'], run_seed=1, max_gens=3)
    assert r['gens'] == 3 and r['fes'] > 381 + 3 * 201 and r['cost'] == r2['cost'] and r['fes'] == r2['fes']
    r = orr.nrlpso_rollout(inst, 10, 20000, A2['qlpso_q'], run_seed=1, max_gens=120)
    assert r['gens'] == 120 and 220 <= r['fes'] <= 100 + 3 * 120
    r = orr.les_rollout(inst, 10, 20000, {'attn': GL['attn'], 'mlp': GL['mlp']}, run_seed=1, max_gens=10)
    assert r['gens'] == 10 and r['fes'] == 16 * 11 and 0 <= r['ret'] < 1
    assert set(orr.ROLLOUTS) >= {'rlepso', 'lde', 'deddqn', 'gleet', 'dedqn', 'rl_hpsde', 'nrlpso', 'les'}
