#!/usr/bin/env python
"""Per-function cost table on one GPU: for every function id of a suite, time the stand-alone evaluation kernel and
one fused optimizer step on B lanes that all run that function (CUDA events, 10 launches after 3 warm-ups).
usage: python scripts/per_function.py [--suit bbob|bbob-noisy] [--dim 30] [--lanes 1184] > profiles/...txt"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metabox_b200 import optimizer as mo, problem as mp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--suit', default='bbob'); ap.add_argument('--dim', type=int, default=30)
ap.add_argument('--lanes', type=int, default=1184); ap.add_argument('--reps', type=int, default=10)
ap.add_argument('--algos', default='eval,rlepso,lde,deddqn')
a = ap.parse_args()
tr, te = mp.BBOB_Dataset.get_datasets(a.suit, a.dim, 5)
probs = sorted(tr.data + te.data, key=lambda p: p.fid)
pack = mp.InstancePack(probs)
dev = pack.device
B, NP, D = a.lanes, 100, a.dim


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.reps


print(f'# {torch.cuda.get_device_name(0)}  suit={a.suit} dim={D} pop={NP} lanes={B}; ms per launch and M evals/s per launch')
print(f'{"fid":>4} {"name":26s} ' + ' '.join(f'{k + "_ms":>10s} {k + "_Mev/s":>12s}' for k in a.algos.split(',')))
g = torch.Generator(device='cpu').manual_seed(0)
for k, p in enumerate(probs):
    lane = np.full(B, k, dtype=np.int32)
    row = []
    for algo in a.algos.split(','):
        if algo == 'eval':
            x = (torch.rand(B, NP, D, generator=g) * 10 - 5).to(dev)
            lanes_t = torch.as_tensor(lane).to(dev)
            out = torch.empty(B, NP, device=dev)
            import ctypes as C
            from metabox_b200 import _lib
            L = _lib.lib()
            fn = lambda: _lib.check(L.mbx_eval(pack.kind, _lib.ptr(pack.headers), _lib.ptr(pack.consts), _lib.ptr(lanes_t), _lib.ptr(x), 0,
                                               _lib.ptr(out), B, NP, D, 1, 1, 0, None, _lib.stream_ptr()), 'mbx_eval')
            n = B * NP
        elif algo == 'rlepso':
            eng = mo.RLEPSO_Engine(pack, lane, 10 ** 9, 50, log_interval=10 ** 7, seed=1); eng.init_population()
            act = torch.rand(B, 35, generator=g).to(dev)
            fn = lambda eng=eng, act=act: eng.update(act)
            n = B * NP
        elif algo == 'gleet':
            eng = mo.GLEET_Engine(pack, lane, 10 ** 9, 50, log_interval=10 ** 7, seed=1); eng.init_population()
            act = torch.rand(B, NP, generator=g).to(dev)
            fn = lambda eng=eng, act=act: eng.update(act)
            n = B * NP
        elif algo == 'lde':
            eng = mo.LDE_Engine(pack, lane, 10 ** 9, 50, log_interval=10 ** 7, seed=1, NP=NP); eng.init_population()
            act = torch.rand(B, 2 * NP, generator=g).to(dev)
            fn = lambda eng=eng, act=act: eng.update(act)
            n = B * NP
        else:
            eng = mo.DE_DDQN_Engine(pack, lane, 10 ** 9, 50, log_interval=10 ** 7, seed=1); eng.init_population()
            act = torch.randint(0, 4, (B,), generator=g).to(torch.int32).to(dev)
            fn = lambda eng=eng, act=act: eng.update(act)
            n = B
        ms = timed(fn)
        row.append(f'{ms:10.4f} {n / ms / 1e3:12.2f}')
    print(f'{p.fid:4d} {str(p):26s} ' + ' '.join(row), flush=True)
