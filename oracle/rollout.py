"""Oracle: complete rollouts (`agent.rollout_episode(PBO_Env(problem, optimizer))`, rlepso_agent.py:294-303)
on the CPU, used as the timed CPU baseline of bench.py and by statistical tests.  TEST INFRASTRUCTURE ONLY."""
import time

import numpy as np
import torch

from . import agents as oag
from . import bbob as ob
from . import rlepso as orl


def rlepso_rollout(inst, dim, maxFEs, weights, run_seed, max_gens=None, n_logpoint=50):
    """One (instance, run) RLEPSO rollout with np.random.seed(run) / torch.manual_seed(run) like the tester
    (tester.py:198; the torch seed is our addition, the reference leaves torch unseeded).
    Returns dict(cost, fes, return, gens, seconds)."""
    np.random.seed(run_seed)
    torch.manual_seed(run_seed)
    cfg = orl.Config(dim=dim, maxFEs=maxFEs, n_logpoint=n_logpoint)
    nc = ob.noise_draw_count(inst)
    ev = lambda x, draws: ob.apply_noise(inst, ob.func_true(inst, x), draws=draws)
    t0 = time.perf_counter()
    S = orl.init_population(cfg, ev, inst.optimum, inst.lb, inst.ub, orl.NumpyTape(cfg.NP, dim, nc))
    feat = np.array([S.fes / maxFEs])
    R, gens, done = 0, 0, False
    while not done and (max_gens is None or gens < max_gens):
        action = oag.rlepso_act(weights, feat.astype(np.float32))
        S, feat, reward, done = orl.update(cfg, S, action, ev, inst.optimum, inst.lb, inst.ub, orl.NumpyTape(cfg.NP, dim, nc))
        R += reward
        gens += 1
    return dict(cost=[float(c) for c in S.curve], fes=S.fes, ret=R, gens=gens, seconds=time.perf_counter() - t0)


def lde_rollout(inst, dim, maxFEs, weights, run_seed, max_gens=None, n_logpoint=50, NP=50):
    """One LDE rollout (lde_agent.py:147-163).  `weights` None -> randomly initialised policy of the right width
    (the shipped policy only fits NP = 50)."""
    from . import lde as ol
    np.random.seed(run_seed)
    torch.manual_seed(run_seed)
    cfg = ol.Config(dim=dim, maxFEs=maxFEs, n_logpoint=n_logpoint, NP=NP)
    nc = ob.noise_draw_count(inst)
    ev = lambda x, draws: ob.apply_noise(inst, ob.func_true(inst, x), draws=draws)
    if weights is None:
        lstm = torch.nn.LSTM(NP + 10, 50, 1); mu_l = torch.nn.Linear(50, 2 * NP); sg_l = torch.nn.Linear(50, 2 * NP)
        weights = {'_PolicyNet__lstm.' + k: v.detach() for k, v in lstm.state_dict().items()}
        weights.update({'_PolicyNet__mu.weight': mu_l.weight.detach(), '_PolicyNet__mu.bias': mu_l.bias.detach(),
                        '_PolicyNet__sigma.weight': sg_l.weight.detach(), '_PolicyNet__sigma.bias': sg_l.bias.detach()})
    t0 = time.perf_counter()
    S, feat = ol.init_population(cfg, ev, inst.optimum, inst.lb, inst.ub, ol.NumpyTape(NP, dim, nc))
    h = torch.zeros(1, 1, 50); c = torch.zeros(1, 1, 50)
    R, gens, done = 0., 0, False
    while not done and (max_gens is None or gens < max_gens):
        mu, sigma, h, c = oag.lde_forward(weights, torch.as_tensor(feat[None, None, :], dtype=torch.float32), h, c)
        action = torch.clip(torch.normal(mu, sigma), 0, 1).reshape(-1).numpy()
        S, feat, reward, done = ol.update(cfg, S, action, ev, inst.optimum, inst.lb, inst.ub, ol.NumpyTape(NP, dim, nc))
        R += float(reward)
        gens += 1
    return dict(cost=[float(c_) for c_ in S.curve], fes=S.fes, ret=R, gens=gens, seconds=time.perf_counter() - t0)


def deddqn_rollout(inst, dim, maxFEs, weights, run_seed, max_gens=None, n_logpoint=50):
    """One DE-DDQN rollout (de_ddqn_agent.py:108-117), greedy actions; `max_gens` bounds the number of STEPS."""
    from . import deddqn as od
    np.random.seed(run_seed)
    cfg = od.Config(dim=dim, maxFEs=maxFEs, n_logpoint=n_logpoint)
    nc = ob.noise_draw_count(inst)
    ev = lambda x, draws: ob.apply_noise(inst, ob.func_true(inst, x), draws=draws)
    t0 = time.perf_counter()
    S, feat = od.init_population(cfg, ev, inst.optimum, inst.lb, inst.ub, od.NumpyTape(cfg.NP, dim, nc))
    fes0 = 0                               # the initial NP evaluations are not part of the per-step rate
    R, steps, done = 0., 0, False
    while not done and (max_gens is None or steps < max_gens):
        q = oag.deddqn_q(weights, np.nan_to_num(feat).astype(np.float32))
        S, feat, reward, done = od.update(cfg, S, int(torch.argmax(q)), ev, inst.optimum, inst.lb, inst.ub,
                                          od.NumpyTape(cfg.NP, dim, nc))
        R += float(reward)
        steps += 1
    return dict(cost=[float(c_) for c_ in S.curve], fes=S.fes - fes0, ret=R, gens=steps, seconds=time.perf_counter() - t0)


def gleet_rollout(inst, dim, maxFEs, weights, run_seed, max_gens=None, n_logpoint=50):
    """One GLEET rollout (gleet_agent.py:293-303).  `weights` = the reference actor's state_dict (attention actor, oracle.agents.gleet_act);
    None -> bench.py's stand-in policy (per-particle MLP 27 -> 64 -> 1, sigmoid; seed 0)."""
    from . import gleet as og
    torch.manual_seed(0)
    if weights is None:
        net = torch.nn.Sequential(torch.nn.Linear(27, 64), torch.nn.Tanh(), torch.nn.Linear(64, 1), torch.nn.Sigmoid(), torch.nn.Flatten(0))
        act = lambda st: net(torch.as_tensor(st, dtype=torch.float32)).numpy()
    else:
        act = lambda st: oag.gleet_act(weights, np.asarray(st, dtype=np.float32))
    torch.manual_seed(run_seed)
    np.random.seed(run_seed)
    cfg = og.Config(dim=dim, maxFEs=maxFEs, n_logpoint=n_logpoint)
    nc = ob.noise_draw_count(inst)
    ev = lambda x, draws: ob.apply_noise(inst, ob.func_true(inst, x), draws=draws)
    t0 = time.perf_counter()
    S, state = og.init_population(cfg, ev, inst.optimum, inst.lb, inst.ub, og.NumpyTape(cfg.NP, dim, nc))
    R, gens, done = 0., 0, False
    with torch.no_grad():
        while not done and (max_gens is None or gens < max_gens):
            action = act(state)
            S, state, reward, done = og.update(cfg, S, action, ev, inst.optimum, inst.lb, inst.ub, og.NumpyTape(cfg.NP, dim, nc))
            R += float(reward)
            gens += 1
    return dict(cost=[float(c_) for c_ in S.curve], fes=S.fes, ret=R, gens=gens, seconds=time.perf_counter() - t0)


def _table_action(q_table, state):
    """Tabular agents (qlpso / rl_hpsde / nrlpso `__get_action`): sample from softmax(Q[state]) with np.random.choice like the reference."""
    e = np.exp(q_table[state])
    return int(np.random.choice(len(e), size=1, p=e / e.sum())[0])


def dedqn_rollout(inst, dim, maxFEs, weights, run_seed, max_gens=None, n_logpoint=50):
    """One DEDQN rollout (dedqn_agent.py:87-96): greedy argmax of the 4 -> 10 -> 10 -> 3 Q-network; `max_gens` bounds the STEPS (200 charged
    evaluations each).  `weights` = the reference state_dict of the Q-network."""
    from . import dedqn as od
    np.random.seed(run_seed)
    cfg = od.Config(dim=dim, maxFEs=maxFEs, n_logpoint=n_logpoint)
    nc = ob.noise_draw_count(inst)
    ev = lambda x, draws: ob.apply_noise(inst, ob.func_true(inst, x), draws=draws)
    w = {('.' + k): torch.as_tensor(np.asarray(v)) for k, v in weights.items()}
    t0 = time.perf_counter()
    S, state = od.init_population(cfg, ev, inst.optimum, inst.lb, inst.ub, od.NumpyTape(cfg.NP, dim, nc))
    R, steps, done = 0., 0, False
    while not done and (max_gens is None or steps < max_gens):
        q = oag._mlp(w, '', torch.as_tensor(np.nan_to_num(state).astype(np.float32)), 3)
        S, state, reward, done = od.update(cfg, S, int(torch.argmax(q)), ev, inst.optimum, inst.lb, inst.ub, od.NumpyTape(cfg.NP, dim, nc))
        R += float(reward); steps += 1
    return dict(cost=[float(c_) for c_ in S.curve], fes=S.fes, ret=R, gens=steps, seconds=time.perf_counter() - t0)


def rl_hpsde_rollout(inst, dim, maxFEs, weights, run_seed, max_gens=None, n_logpoint=50):
    """One RL_HPSDE rollout (rl_hpsde_agent.py:62-72); `weights` = the [4, 4] Q-table."""
    from . import rl_hpsde as oh
    np.random.seed(run_seed)
    cfg = oh.Config(dim=dim, maxFEs=maxFEs, n_logpoint=n_logpoint)
    nc = ob.noise_draw_count(inst)
    ev = lambda x, draws: ob.apply_noise(inst, ob.func_true(inst, x), draws=draws)
    t0 = time.perf_counter()
    S, state = oh.init_population(cfg, ev, inst.optimum, inst.lb, inst.ub, oh.NumpyTape(dim, nc))
    R, gens, done = 0., 0, False
    while not done and (max_gens is None or gens < max_gens):
        S, state, reward, done = oh.update(cfg, S, _table_action(np.asarray(weights), state), ev, inst.optimum, inst.lb, inst.ub, oh.NumpyTape(dim, nc))
        R += float(reward); gens += 1
    return dict(cost=[float(c_) for c_ in S.curve], fes=S.fes, ret=R, gens=gens, seconds=time.perf_counter() - t0)


def nrlpso_rollout(inst, dim, maxFEs, weights, run_seed, max_gens=None, n_logpoint=50):
    """One NRLPSO rollout; `weights` = a [4, 4] Q-table; `max_gens` bounds the single-particle STEPS."""
    from . import nrlpso as on
    np.random.seed(run_seed)
    cfg = on.Config(dim=dim, maxFEs=maxFEs, n_logpoint=n_logpoint)
    nc = ob.noise_draw_count(inst)
    ev = lambda x, draws: ob.apply_noise(inst, ob.func_true(inst, x), draws=draws)
    t0 = time.perf_counter()
    S, state = on.init_population(cfg, ev, inst.optimum, inst.lb, inst.ub, on.NumpyTape(dim, nc))
    R, steps, done = 0., 0, False
    while not done and (max_gens is None or steps < max_gens):
        S, state, reward, done = on.update(cfg, S, _table_action(np.asarray(weights), state), ev, inst.optimum, inst.lb, inst.ub, on.NumpyTape(dim, nc))
        R += float(reward); steps += 1
    return dict(cost=[float(c_) for c_ in S.curve], fes=S.fes, ret=R, gens=steps, seconds=time.perf_counter() - t0)


def les_rollout(inst, dim, maxFEs, weights, run_seed, max_gens=None, n_logpoint=50):
    """One LES episode; `weights` = {'attn': [68], 'mlp': [178]} (noise-free functions: the oracle's LES evaluates raw function values)."""
    from . import les as ol
    np.random.seed(run_seed)
    ev = lambda x, draws: ob.func_true(inst, x)
    t0 = time.perf_counter()
    S = ol.init_population(dim, maxFEs, ev, inst.lb, inst.ub, n_logpoint)
    gens, y0 = 0, None
    while not S.done and (max_gens is None or gens < max_gens):
        S = ol.generation(S, weights['attn'], weights['mlp'], dim, maxFEs, maxFEs // n_logpoint, ev, inst.optimum, inst.lb, inst.ub)
        gens += 1
        y0 = S.gbest if y0 is None else y0
    return dict(cost=[float(c_) for c_ in S.curve], fes=S.fes, ret=float((y0 - S.gbest) / y0) if y0 else 0.0, gens=gens, seconds=time.perf_counter() - t0)


ROLLOUTS = {'rlepso': rlepso_rollout, 'lde': lde_rollout, 'deddqn': deddqn_rollout, 'gleet': gleet_rollout, 'dedqn': dedqn_rollout,
            'rl_hpsde': rl_hpsde_rollout, 'nrlpso': nrlpso_rollout, 'les': les_rollout}


_SUITES = {}


def _worker(args):
    torch.set_num_threads(1)
    fid, dim, maxFEs, weights, seed, max_gens, suit, algo, kw = args
    if (suit, dim) not in _SUITES:                       # instances are built once per worker process
        tr, te = ob.make_suite(suit, dim, 5.0)
        _SUITES[(suit, dim)] = {i.fid: i for i in tr + te}
    inst = _SUITES[(suit, dim)][fid]
    r = ROLLOUTS[algo](inst, dim, maxFEs, weights, seed, max_gens, **kw)
    return r['fes'], r['seconds']


def make_pool(n_procs):
    """A fork pool of single-threaded workers (None for n_procs <= 1); create it once, outside the timed region."""
    import multiprocessing as mp
    return mp.get_context('fork').Pool(n_procs) if n_procs > 1 else None


def timed_sample(dim, maxFEs, weights, fids, max_gens, n_procs, suit='bbob', seed0=0, algo='rlepso', runs=1, pool=None, **kw):
    """Run one bounded sample: each function id in `fids` gets `runs` runs truncated at `max_gens` generations (steps for
    DE-DDQN), spread over a pool of `n_procs` single-threaded workers (heaviest jobs are not known in advance: chunksize 1).
    -> (total fes, wall seconds of the map itself when `pool` is given, else including the pool start-up)."""
    jobs = [(f, dim, maxFEs, weights, seed0 + 1000 * r + k, max_gens, suit, algo, kw) for r in range(runs) for k, f in enumerate(fids)]
    own = pool is None and n_procs > 1
    t0 = time.perf_counter()
    if own:
        pool = make_pool(n_procs)
    if pool is None:
        res = [_worker(j) for j in jobs]
    else:
        res = pool.map(_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    if own:
        pool.close(); pool.join()
    return sum(r[0] for r in res), wall
