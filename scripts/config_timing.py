#!/usr/bin/env python
"""Throughput of the BASELINE.json parity configurations C2 (RLEPSO, bbob 24 fns, D10, 64 runs) and C5 (RLEPSO, 280 protein
instances x 51 runs, D12, maxFEs 1000) on one GPU -- documentation numbers, not bench lines (the metric is quoted on D30)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from metabox_b200 import agent as ma, problem as mp, sweep   # noqa: E402

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'agents.npz'))
w = {k[len('rlepso/'):]: G[k] for k in G.files if k.startswith('rlepso/')}
actor = ma.load_reference_weights(ma.RLEPSO_Actor(), w).cuda().eval()


def run(name, pack, n_inst, runs, maxFEs, n_log, reps=3):
    inst, _ = sweep.lane_table(n_inst, runs)
    sw = sweep.StripedSweep(lambda ln, lb: sweep.RLEPSO_Sweep(pack, ln, actor, maxFEs, n_log, seed=1, lane_base=lb), inst, 0, 3)
    sw.run(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fes = 0
    for _ in range(reps):
        fes += int(sw.run()['fes'].sum().item())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f'{name}: {len(inst)} lanes, {fes / reps / 1e6:.2f} M evals per sweep, {dt / reps * 1e3:.1f} ms per sweep, {fes / dt / 1e6:.1f} M evals/s')


tr, te = mp.BBOB_Dataset.get_datasets('bbob', 10, 5)
run('C2 RLEPSO bbob D10 x 64 runs (maxFEs 20000)', mp.InstancePack(tr.data + te.data), 24, 64, 20000, 50)
run('C2 at 1024 runs', mp.InstancePack(tr.data + te.data), 24, 1024, 20000, 50)
probs = [mp.synthetic_protein(1000 + k) for k in range(280)]
run('C5 RLEPSO protein 280 x 51 runs (maxFEs 1000)', mp.InstancePack(probs), 280, 51, 1000, 5)
