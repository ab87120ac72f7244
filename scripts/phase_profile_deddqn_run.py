#!/usr/bin/env python
"""Persistent DE-DDQN kernel (mbx_deddqn_run): microseconds per environment step and, with the instrumented build
(MBX_LIB=.../libmbx_prof.so), the share of thread-0 cycles per phase.   python scripts/phase_profile_deddqn_run.py [runs] [n_steps]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from metabox_b200 import _lib, agent as ma, optimizer as mo, problem as mp   # noqa: E402

L = _lib.lib()
prof = hasattr(L, 'mbx_prof_read_deddqn_run')
dev = torch.device('cuda', 0)
tr, te = mp.BBOB_Dataset.get_datasets('bbob', 30, 5, difficulty='difficult')
pack = mp.InstancePack(te.data, dev)
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'agents.npz'))
net = ma.DE_DDQN_FusedNet(ma.load_reference_weights(ma.DE_DDQN_Net(), {k[7:]: G[k] for k in G.files if k.startswith('deddqn/')}).to(dev).eval())
names = {0: 'gen boundary', 1: 'mutation', 2: 'median rank', 3: 'trial fitness', 4: 'credit/window', 5: 'selection+log', 6: 'features',
         8: 'loop-top barrier', 9: 'Q-network', 10: 'env step total'}
buf = (C.c_ulonglong * 16)()
runs_list = [int(sys.argv[1])] if len(sys.argv) > 1 else [16, 64, 256]
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
for runs in runs_list:
    lanes = np.repeat(np.arange(18, dtype=np.int32), runs)
    eng = mo.DE_DDQN_Engine(pack, lanes, 60000, 50, seed=7)
    eng.init_population()
    ret = torch.zeros(len(lanes), device=dev)
    eng.run_steps(net.packed, ret, 150)
    torch.cuda.synchronize()
    if prof:
        L.mbx_prof_read_deddqn_run(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.run_steps(net.packed, ret, n_steps)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    line = f'lanes={len(lanes):5d} n_steps={n_steps}: {ms:8.2f} ms per launch = {1000 * ms / n_steps:7.2f} us per step of all lanes = {len(lanes) * n_steps / ms / 1e3:7.2f} M evals/s'
    if prof:
        L.mbx_prof_read_deddqn_run(buf, 1)
        v = np.array(list(buf), dtype=np.float64)
        ctas = -(-len(lanes) // L.mbx_deddqn_run_lanes_per_cta())
        tot = v[8] + v[9] + v[10]
        line += f' | cycles per CTA-step {tot / n_steps / ctas:8.0f} | ' + ' | '.join(f'{names[k]} {v[k] / n_steps / ctas:7.0f}' for k in sorted(names))
    print(line)
