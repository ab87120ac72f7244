#!/usr/bin/env python
"""Sweep throughput of the two COMPOSED backbones (DEDQN, RL_HPSDE: library kernels + batched torch ops, eager) in charged evaluations per
second, with the share of the step spent inside the library kernels' callers -- the numbers that decide whether a fused step kernel is
worth writing for them.
    python scripts/composed_timing.py [--dim 30] [--runs 64]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from metabox_b200 import agent as ma, problem as mp, sweep   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--dim', type=int, default=30)
ap.add_argument('--runs', type=int, default=64)
a = ap.parse_args()
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')
tr, te = mp.BBOB_Dataset.get_datasets('bbob', a.dim, 5)
probs = tr.data + te.data
pack = mp.InstancePack(probs)
inst, _ = sweep.lane_table(len(probs), a.runs)
maxFEs = 2000 * a.dim
G = np.load(os.path.join(GOLD, 'agents2.npz'))
GL = np.load(os.path.join(GOLD, 'les.npz'))                  # seeded LES network parameters (the backbone's action)
dq = ma.load_reference_weights(ma.DEDQN_Net(), {k[6:]: G[k] for k in G.files if k.startswith('dedqn/')}).cuda().eval()
qt = ma.RL_HPSDE_QTable(G['rl_hpsde_q']).cuda()           # the shipped agent's Q-table
nq = ma.NRLPSO_QTable(G['qlpso_q']).cuda()                  # (no NRLPSO agent is shipped; any 4 x 4 table drives the backbone)
small = sweep.lane_table(len(probs), 8)[0]                  # NRLPSO: [lanes, 100, 100, D] float64 distance temporaries, one particle per step
for name, make in (('DEDQN', lambda: sweep.DEDQN_Sweep(pack, inst, dq, maxFEs, 50, seed=1)),
                   ('RL_HPSDE', lambda: sweep.RL_HPSDE_Sweep(pack, inst, ma.fuse(qt), maxFEs, 50, seed=1)),
                   ('NRLPSO', lambda: sweep.NRLPSO_Sweep(pack, small, ma.fuse(nq), min(maxFEs, 3000), 50, seed=1)),
                   ('LES', lambda: sweep.LES_Sweep(pack, inst, {'attn': GL['attn'], 'mlp': GL['mlp']}, maxFEs, 50, seed=1))):
    sw = make()
    sw.run()                                               # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = sw.run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    fes = int(res['fes'].sum())
    print(f'{name:9s} dim {a.dim}: {sw.engine.B} lanes x {res["steps"]} steps, {fes / 1e6:8.1f} M charged evaluations in {ms:8.1f} ms = '
          f'{fes / ms / 1e3:8.1f} M evals/s ({ms / max(1, res["steps"]):6.2f} ms per step, composed / eager)')
