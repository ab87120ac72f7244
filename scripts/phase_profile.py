#!/usr/bin/env python
"""Per-phase share of the fused RLEPSO step (clock64 deltas of thread 0 of every CTA, summed over CTAs).
Needs the instrumented build:  python -m metabox_b200.build --profile ; MBX_LIB=metabox_b200/csrc/libmbx_prof.so
usage: MBX_LIB=... python scripts/phase_profile.py [--runs 64] [--steps 40] [--fids 1,15,16,21] """
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from metabox_b200 import _lib, optimizer as mo, problem as mp   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--runs', type=int, default=64)
ap.add_argument('--steps', type=int, default=40)
ap.add_argument('--dim', type=int, default=30)
ap.add_argument('--warm', type=int, default=3, help='generations before the measured ones (late generations = converged swarms)')
ap.add_argument('--fids', default='all,1,3,15,16,17,21,23')
a = ap.parse_args()
L = _lib.lib()
assert hasattr(L, 'mbx_prof_read'), 'load the instrumented library through MBX_LIB'
dev = torch.device('cuda', 0)
tr, te = mp.BBOB_Dataset.get_datasets('bbob', a.dim, 5)
probs = tr.data + te.data
pack = mp.InstancePack(probs, dev)
names = ['P0 stage+scalars', 'P0b rank+sorted copy', 'P1 FDR scan', 'P2 velocity/position', 'P3 fitness', 'P4-5 pbest/gbest/reinit']
buf = (C.c_ulonglong * 16)()
print(f'# {torch.cuda.get_device_name(0)} dim={a.dim} runs/fn={a.runs} steps={a.steps} after {a.warm} warm generations; share of CTA cycles per phase, ms per launch')
for sel in a.fids.split(','):
    ids = list(range(len(probs))) if sel == 'all' else [k for k, p in enumerate(probs) if pack.h_headers['fid'][k] == int(sel)]
    lanes = np.repeat(np.asarray(ids, dtype=np.int32), a.runs * (1 if sel == 'all' else 8))
    eng = mo.RLEPSO_Engine(pack, lanes, 2000 * a.dim, 50, seed=7)
    eng.init_population()
    g = torch.Generator(device='cpu').manual_seed(0)
    act = torch.rand(len(lanes), 35, generator=g).to(dev)
    for _ in range(a.warm):
        eng.update(act)
    torch.cuda.synchronize()
    L.mbx_prof_read(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        eng.update(act)
    e1.record(); torch.cuda.synchronize()
    L.mbx_prof_read(buf, 1)
    v = np.array(list(buf)[:6], dtype=np.float64)
    fb = buf[6] / max(1, buf[7])
    ms = e0.elapsed_time(e1) / a.steps
    print(f'fid={sel:>4} lanes={len(lanes):5d} {ms:8.4f} ms/launch  {len(lanes) * 100 / ms / 1e3:8.1f} Mev/s | ' +
          ' | '.join(f'{n.split()[0]} {100 * x / v.sum():5.1f}%' for n, x in zip(names, v)) + f' | FDR exact-pass items {100 * fb:.3f}%' +
          (f' | Gallagher scan {100 * buf[8] / v.sum():.1f}% exact {100 * buf[9] / v.sum():.1f}% f64-fallbacks/launch {buf[10] / a.steps:.1f}' if buf[8] else ''))
