#!/usr/bin/env python
"""Device time of the two fused policy kernels against their torch modules on benchmark-sized batches (CUDA events on the launching stream,
warm-up, L2 flushed between repetitions), and a bounded launch sequence for `ncu`:
    python scripts/policy_timing.py                       # table: torch module vs fused kernel, ms per call
    ncu --set full --clock-control none -k regex:gleet_policy -s 3 -c 1 python scripts/policy_timing.py --only gleet --reps 6"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from metabox_b200 import agent as ma   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--only', default='both', choices=['both', 'gleet', 'deddqn', 'rlepso'])
ap.add_argument('--reps', type=int, default=20)
a = ap.parse_args()
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')            # > 126 MB L2


def timed(fn, reps):
    for _ in range(3):
        fn()
    ms = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


if a.only in ('both', 'gleet'):
    G = np.load(os.path.join(GOLD, 'agents2.npz'))
    actor = ma.load_reference_weights(ma.GLEET_Actor(), {k[6:]: G[k] for k in G.files if k.startswith('gleet/')}).cuda().eval()
    fused = ma.GLEET_FusedActor(actor)
    for B in (2048, 6144):
        x = torch.rand(B, 100, 27, device='cuda'); nz = torch.randn(B, 100, device='cuda')
        t_torch, t_fused = timed(lambda: actor(x, noise=nz), a.reps), timed(lambda: fused(x, noise=nz), a.reps)
        err = float((actor(x, noise=nz) - fused(x, noise=nz)).abs().max())
        print(f'gleet policy  B={B:5d} NP=100: torch module {t_torch:8.3f} ms   fused kernel {t_fused:8.3f} ms   x{t_torch / t_fused:6.1f}   max |diff| {err:.2e}   '
              f'{B * 100 * 112 / t_fused / 1e6:7.1f} GB/s algorithmic (108 B state + 4 B action per particle)')
if a.only in ('both', 'deddqn'):
    G = np.load(os.path.join(GOLD, 'agents.npz'))
    net = ma.load_reference_weights(ma.DE_DDQN_Net(), {k[7:]: G[k] for k in G.files if k.startswith('deddqn/')}).cuda().eval()
    fused = ma.DE_DDQN_FusedNet(net)
    for B in (768, 4608):
        x = torch.rand(B, 99, device='cuda')
        ref = lambda: torch.argmax(net(torch.nan_to_num(x, nan=0.0, posinf=0.0, neginf=0.0)), dim=-1).to(torch.int32)
        t_torch, t_fused = timed(ref, a.reps), timed(lambda: fused(x), a.reps)
        agree = float((ref() == fused(x)).float().mean())
        print(f'deddqn policy B={B:5d}:        torch module {t_torch:8.3f} ms   fused kernel {t_fused:8.3f} ms   x{t_torch / t_fused:6.1f}   actions equal {agree:.4f}')
if a.only in ('both', 'rlepso'):
    G = np.load(os.path.join(GOLD, 'agents.npz'))
    actor = ma.load_reference_weights(ma.RLEPSO_Actor(), {k[7:]: G[k] for k in G.files if k.startswith('rlepso/')}).cuda().eval()
    fused = ma.RLEPSO_FusedActor(actor)
    for B in (2048, 6144):
        x = torch.rand(B, 1, device='cuda'); nz = torch.randn(B, 35, device='cuda')
        t_torch, t_fused = timed(lambda: actor(x, noise=nz), a.reps), timed(lambda: fused(x, noise=nz), a.reps)
        err = float((actor(x, noise=nz) - fused(x, noise=nz)).abs().max())
        print(f'rlepso policy B={B:5d}:        torch module {t_torch:8.3f} ms   fused kernel {t_fused:8.3f} ms   x{t_torch / t_fused:6.1f}   max |diff| {err:.2e}')
