#!/usr/bin/env python
"""T2 of the SINGLE-PROBLEM drop-in path: what a MetaBox user gets by swapping the two imports of INTEGRATION.md section 2 and changing
nothing else -- `agent.rollout_episode(PBO_Env(problem, <Algo>_Optimizer(config)))` one problem and one run at a time (B = 1: one kernel
launch + one host synchronisation per `env.step`), timed like `Tester.test` does (src/tester.py:196-222: mean wall-clock milliseconds per
run over the runs of the first test problem), next to the reference's own published T2 for that setting
(post_processed_data/wall_time/synthetic-easy.md:4; hardware not stated there).

    python scripts/dropin_timing.py [--runs 51]

The batched sweeps (bench.py) are where the throughput comes from; this is the latency a drop-in user sees."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from metabox_b200 import agent as ma, environment as me, optimizer as mo, problem as mp   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--runs', type=int, default=51)
ap.add_argument('--deddqn-runs', type=int, default=5)
a = ap.parse_args()
dev = torch.device('cuda', 0)
G = np.load(os.path.join(ROOT, 'tests', 'golden', 'agents.npz'))
w = lambda pre: {k[len(pre):]: G[k] for k in G.files if k.startswith(pre)}
PUBLISHED = {'RLEPSO_Agent': 136.34, 'LDE_Agent': 292.33, 'DE_DDQN_Agent': 17509.66}      # synthetic-easy.md:4 (ms per run, 10-D)
dim = 10
cfg = lambda: argparse.Namespace(dim=dim, maxFEs=2000 * dim, n_logpoint=50, log_interval=2000 * dim // 50)
tr, te = mp.BBOB_Dataset.get_datasets('bbob', dim, 5)
problem = te.data[0]                                           # the first test problem, like tester.py:197
mp.InstancePack([problem], dev)


def rollout_rlepso(env, actor):
    state, done, R = env.reset(), False, 0
    while not done:
        action = actor(torch.as_tensor(np.asarray(state), dtype=torch.float32, device=dev).reshape(1, 1))[0].cpu().numpy()
        state, r, done = env.step(action)
        R += r
    return env.optimizer.fes


def rollout_lde(env, net):
    state, done = env.reset(), False
    h = torch.zeros(1, net.cell, device=dev); c = torch.zeros(1, net.cell, device=dev)
    while not done:
        action, h, c = net.step(torch.as_tensor(np.asarray(state), dtype=torch.float32, device=dev).reshape(1, -1), h, c)
        state, r, done = env.step(action.cpu().numpy())
    return env.optimizer.fes


def rollout_deddqn(env, net):
    state, done = env.reset(), False
    while not done:
        q = net(torch.nan_to_num(torch.as_tensor(np.asarray(state), dtype=torch.float32, device=dev).reshape(1, -1)))
        state, r, done = env.step(int(torch.argmax(q)))
    return env.optimizer.fes


cases = [('RLEPSO_Agent', mo.RLEPSO_Optimizer, ma.load_reference_weights(ma.RLEPSO_Actor(), w('rlepso/')), rollout_rlepso, a.runs),
         ('LDE_Agent', mo.LDE_Optimizer, ma.load_reference_weights(ma.LDE_PolicyNet(), w('lde/')), rollout_lde, a.runs),
         ('DE_DDQN_Agent', mo.DE_DDQN_Optimizer, ma.load_reference_weights(ma.DE_DDQN_Net(), w('deddqn/')), rollout_deddqn, a.deddqn_runs)]
print(f'# {torch.cuda.get_device_name(0)}; problem {problem} dim {dim}; mean wall-clock ms per run (tester.py T2), single-problem drop-in path (B = 1)')
with torch.no_grad():
    for name, Opt, net, roll, runs in cases:
        net = net.to(dev).eval()
        env = me.PBO_Env(problem, Opt(cfg()))
        np.random.seed(12345); roll(env, net)                  # warm-up (kernel attributes, cuBLAS handles)
        torch.cuda.synchronize()
        ms, fes = [], []
        for run in range(runs):
            np.random.seed(run)                                # tester.py:198
            torch.manual_seed(run)
            problem.reset()
            t0 = time.perf_counter()
            fes.append(roll(env, net))
            torch.cuda.synchronize()
            ms.append(1000 * (time.perf_counter() - t0))
        print(f'{name:14s} T2 = {np.mean(ms):10.2f} ms per run over {runs} runs (mean fes {np.mean(fes):8.0f}; {1000 * np.mean(ms) / max(1, np.mean(fes) / {"RLEPSO_Agent": 100, "LDE_Agent": 50, "DE_DDQN_Agent": 1}[name]):7.1f} us per env.step)'
              f'   reference, published: {PUBLISHED[name]:10.2f} ms   ratio {PUBLISHED[name] / np.mean(ms):6.2f}x')
