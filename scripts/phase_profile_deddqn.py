#!/usr/bin/env python
"""Per-phase share of the fused DE-DDQN step (instrumented build, see scripts/phase_profile.py)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from metabox_b200 import _lib, optimizer as mo, problem as mp   # noqa: E402

L = _lib.lib()
assert hasattr(L, 'mbx_prof_read_deddqn'), 'load the instrumented library through MBX_LIB'
dev = torch.device('cuda', 0)
tr, te = mp.BBOB_Dataset.get_datasets('bbob', 30, 5, difficulty='difficult')
pack = mp.InstancePack(te.data, dev)
names = ['stage+load+gen boundary', 'mutation', 'median rank', 'trial fitness', 'thread-0 credit/window', 'selection+log', 'features']
buf = (C.c_ulonglong * 16)()
for runs in (64, 256):
    lanes = np.repeat(np.arange(18, dtype=np.int32), runs)
    eng = mo.DE_DDQN_Engine(pack, lanes, 60000, 50, seed=7)
    eng.init_population()
    g = torch.Generator(device='cpu').manual_seed(0)
    act = torch.randint(0, 4, (len(lanes),), generator=g).to(torch.int32).to(dev)
    for _ in range(150):
        eng.update(act)
    torch.cuda.synchronize()
    L.mbx_prof_read_deddqn(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        eng.update(act)
    e1.record(); torch.cuda.synchronize()
    L.mbx_prof_read_deddqn(buf, 1)
    v = np.array(list(buf)[:7], dtype=np.float64)
    print(f'lanes={len(lanes)} {e0.elapsed_time(e1) / 200 * 1000:7.1f} us/step; cycles per CTA-step {v.sum() / 200 / len(lanes):8.0f} | ' +
          ' | '.join(f'{n} {100 * x / v.sum():4.1f}%' for n, x in zip(names, v)))
