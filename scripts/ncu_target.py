#!/usr/bin/env python
"""Small, bounded launch sequences for `ncu` captures of one step kernel (bench.py runs whole sweeps, which is too long under
the profiler for the one-evaluation-per-step backbones):
    ncu --set full --clock-control none -k regex:<kernel> -s 20 -c 1 python scripts/ncu_target.py --algo gleet"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from metabox_b200 import optimizer as mo, problem as mp   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--algo', default='gleet', choices=['gleet', 'deddqn', 'rl_pso', 'rlepso-protein', 'lde'])
ap.add_argument('--runs', type=int, default=256)
ap.add_argument('--steps', type=int, default=30)
a = ap.parse_args()
g = torch.Generator(device='cpu').manual_seed(0)
if a.algo == 'rlepso-protein':
    probs = [mp.synthetic_protein(1000 + k) for k in range(280)]
    pack = mp.InstancePack(probs)
    eng = mo.RLEPSO_Engine(pack, np.repeat(np.arange(280), 51), 10 ** 6, 5, log_interval=10 ** 5, seed=1)
    act = lambda: torch.rand(eng.B, 35, generator=g).cuda()
else:
    tr, te = mp.BBOB_Dataset.get_datasets('bbob-noisy' if a.algo == 'lde' else 'bbob', 30, 5,
                                          difficulty='difficult' if a.algo == 'deddqn' else 'easy')
    probs = te.data if a.algo == 'deddqn' else tr.data + te.data
    pack = mp.InstancePack(probs)
    lanes = np.repeat(np.arange(len(probs)), a.runs)
    if a.algo == 'gleet':
        eng = mo.GLEET_Engine(pack, lanes, 60000, 50, seed=1); act = lambda: torch.rand(eng.B, 100, generator=g).cuda()
    elif a.algo == 'deddqn':
        eng = mo.DE_DDQN_Engine(pack, lanes, 60000, 50, seed=1); act = lambda: torch.randint(0, 4, (eng.B,), generator=g).to(torch.int32).cuda()
    elif a.algo == 'rl_pso':
        eng = mo.RL_PSO_Engine(pack, lanes, 60000, 50, seed=1); act = lambda: torch.rand(eng.B, generator=g).cuda()
    else:
        eng = mo.LDE_Engine(pack, lanes, 60000, 50, seed=1, NP=100); act = lambda: torch.rand(eng.B, 200, generator=g).cuda()
eng.init_population()
acts = [act() for _ in range(a.steps)]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
eng.update(acts[0])
torch.cuda.synchronize()
e0.record()
for k in range(1, a.steps):
    eng.update(acts[k])
e1.record()
torch.cuda.synchronize()
print(a.algo, 'lanes', eng.B, 'steps', a.steps, 'fes[0]', int(eng.fes[0]),
      'ms/step (incl. launch overhead, not valid under ncu) %.4f' % (e0.elapsed_time(e1) / max(1, a.steps - 1)))
