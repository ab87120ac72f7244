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


def _worker(args):
    torch.set_num_threads(1)
    fid, dim, maxFEs, weights, seed, max_gens, suit = args
    tr, te = ob.make_suite(suit, dim, 5.0)
    inst = [i for i in tr + te if i.fid == fid][0]
    r = rlepso_rollout(inst, dim, maxFEs, weights, seed, max_gens)
    return r['fes'], r['seconds']


def timed_sample(dim, maxFEs, weights, fids, max_gens, n_procs, suit='bbob', seed0=0):
    """Run one bounded sample: each function id in `fids` gets one run truncated at `max_gens` generations,
    spread over a pool of `n_procs` single-threaded workers.  -> (total fes, wall seconds)."""
    import multiprocessing as mp
    jobs = [(f, dim, maxFEs, weights, seed0 + k, max_gens, suit) for k, f in enumerate(fids)]
    t0 = time.perf_counter()
    if n_procs <= 1:
        res = [_worker(j) for j in jobs]
    else:
        with mp.get_context('fork').Pool(n_procs) as pool:
            res = pool.map(_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    return sum(r[0] for r in res), wall
