#!/usr/bin/env python
"""Device time and algorithmic-byte roofline of every library kernel that has no line in bench.py: the stand-alone evaluation kernel per
function, Random_search, the one-particle-per-step backbones (RL_PSO, QLPSO), the landscape features, the progressive walk and the five
fused policy kernels.  CUDA events on the launching stream, 3 warm-ups, inputs larger than L2 or L2 flushed between repetitions (said per
line), median of `--reps` launches.

    python scripts/kernel_rooflines.py [--only eval,rs,...] [--reps 10]          # table
    ncu --set full --clock-control none --import-source on -k regex:'<kernel>' -c 1 -s 3 python scripts/kernel_rooflines.py --only <name> --reps 2

Algorithmic bytes per unit: eval 4 D + 4 per evaluation; Random_search the same (positions are generated in the kernel: 4 B
written per evaluation would undercount -- the per-evaluation figure of the eval path is used); one-particle steps: rows read + row written +
state vector; policies: state in + action out (weights are L2 resident and amortised)."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from metabox_b200 import _lib, agent as ma, optimizer as mo, problem as mp   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--only', default='eval,rs,rl_pso,qlpso,landscape,walk,policies')
ap.add_argument('--reps', type=int, default=10)
ap.add_argument('--lanes', type=int, default=16384, help='lanes of the eval launches (16384 x 100 x 30 x 4 B = 197 MB of positions > 126 MB L2)')
ap.add_argument('--fids', default='1,5,8,20,2,15,16,21,23')
a = ap.parse_args()
only = set(a.only.split(','))
dev = torch.device('cuda', 0)
L = _lib.lib()
try:
    PEAK = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    PEAK = 6650.0
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)            # > 126 MB L2
GOLD = os.path.join(ROOT, 'tests', 'golden')


def timed(fn, flush_l2=False):
    for _ in range(3):
        fn()
    ms = []
    for _ in range(a.reps):
        if flush_l2:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def line(name, ms, units, bytes_per_unit, note):
    gbs = units * bytes_per_unit / (ms / 1e3) / 1e9
    print(f'{name:44s} {ms:9.4f} ms  {units / ms / 1e3:10.1f} M units/s  {gbs:8.1f} GB/s algorithmic = {100 * gbs / PEAK:5.1f} % of {PEAK:.0f} GB/s   {note}')


print(f'# {torch.cuda.get_device_name(0)}; peak = measured HBM copy bandwidth (MEASURED_PEAKS.json) or the B200_PROFILING.md fallback')
D, NP = 30, 100
tr, te = mp.BBOB_Dataset.get_datasets('bbob', D, 5)
probs = sorted(tr.data + te.data, key=lambda p: p.fid)
pack = mp.InstancePack(probs, dev)
g = torch.Generator(device='cpu').manual_seed(0)

if 'eval' in only:
    B = a.lanes
    x = (torch.rand(B, NP, D, generator=g) * 10 - 5).to(dev)
    out = torch.empty(B, NP, device=dev)
    for fid in [int(f) for f in a.fids.split(',')]:
        k = [p.fid for p in probs].index(fid)
        lanes_t = torch.full((B,), k, dtype=torch.int32, device=dev)
        kind = pack.kind | (_lib.EVAL_ALL_STREAMING if fid in (5, 8, 20) else _lib.EVAL_NONE_STREAMING)      # every lane runs function fid
        fn = lambda kind=kind, lanes_t=lanes_t: _lib.check(L.mbx_eval(kind, _lib.ptr(pack.headers), _lib.ptr(pack.consts), _lib.ptr(lanes_t), _lib.ptr(x), 0,
                                           _lib.ptr(out), B, NP, D, 1, 1, 0, None, _lib.stream_ptr()), 'mbx_eval')
        line(f'eval_kernel<30,BBOB,f32> F{fid} {str(probs[k])[:18]}', timed(fn), B * NP, 4 * D + 4, f'{B} lanes x {NP} candidates, inputs {x.numel() * 4 / 1e6:.0f} MB > L2')

if 'rs' in only:
    import argparse as _ap
    cfg = _ap.Namespace(dim=D, maxFEs=2000 * D, n_logpoint=50, log_interval=2000 * D // 50)
    opt = mo.Random_search(cfg)
    for fid in (1, 15):
        k = [p.fid for p in probs].index(fid)
        lane = np.full(4096, k, dtype=np.int32)
        ms = timed(lambda: opt.run_lanes(pack, lane, seed=3))
        line(f'random_search_kernel<30,BBOB> F{fid}', ms, 4096 * 60000, 4 * D + 4, '4096 lanes x 60000 evaluations, whole episodes in one launch (incl. result D2H)')

if 'rl_pso' in only or 'qlpso' in only:
    lanes = np.repeat(np.arange(len(probs)), 256)
    if 'rl_pso' in only:
        eng = mo.RL_PSO_Engine(pack, lanes, 60000, 50, seed=1); eng.init_population()
        act = torch.rand(eng.B, generator=g).to(dev)
        line('rlpso_step_kernel<30,BBOB>', timed(lambda: eng.update(act), True), eng.B, 5 * 4 * D + 8 * D + 16, f'{eng.B} lanes, one particle per step (L2 flushed)')
    if 'qlpso' in only:
        eng = mo.QLPSO_Engine(pack, lanes, 60000, 50, seed=1); eng.init_population()
        act = torch.randint(0, 4, (eng.B,), generator=g).to(torch.int32).to(dev)
        line('qlpso_step_kernel<30,BBOB>', timed(lambda: eng.update(act), True), eng.B, 4 * 4 * D + 4 * NP * D + 16, f'{eng.B} lanes, one particle per step + swarm diversity (L2 flushed)')

if 'landscape' in only:
    from metabox_b200 import landscape as ml
    Bn, S = 8192, 201
    smp = torch.rand(Bn, S, D, dtype=torch.float64, generator=g).to(dev); cst = torch.rand(Bn, S, dtype=torch.float64, generator=g).to(dev)
    ref = torch.rand(Bn, D, dtype=torch.float64, generator=g).to(dev)
    line('landscape_kernel', timed(lambda: ml.features(smp, cst, ref)), Bn, 8 * (S * D + S + D) + 48, f'{Bn} walks of {S} points x {D} dims (float64), inputs {smp.numel() * 8 / 1e6:.0f} MB > L2')

if 'walk' in only:
    from metabox_b200 import rl_hpsde as mh
    Bn, S = 8192, 200
    zone = torch.rand(Bn, D, dtype=torch.float64, generator=g).to(dev); ru = torch.rand(Bn, S, dtype=torch.float64, generator=g).to(dev)
    rD = torch.randint(0, D, (Bn, S), generator=g).to(torch.int32).to(dev); su = torch.rand(Bn, S, D, dtype=torch.float64, generator=g).to(dev)
    out64 = torch.empty(Bn, S + 1, D, dtype=torch.float64, device=dev); out32 = torch.empty(Bn, S + 1, D, dtype=torch.float32, device=dev)
    fn = lambda: _lib.check(L.mbx_progressive_walk(_lib.ptr(zone), _lib.ptr(ru), _lib.ptr(rD), _lib.ptr(su), -5.0, 5.0, 10.0, Bn, D, S,
                                                   _lib.ptr(out64), _lib.ptr(out32), _lib.stream_ptr()), 'mbx_progressive_walk')
    line('progressive_walk_kernel', timed(fn), Bn, 8 * S * D + 12 * (S + 1) * D + 12 * S, f'{Bn} walks x {S} steps x {D} dims')

if 'policies' in only:
    G1, G2 = np.load(os.path.join(GOLD, 'agents.npz')), np.load(os.path.join(GOLD, 'agents2.npz'))
    w = lambda G, pre: {k[len(pre):]: G[k] for k in G.files if k.startswith(pre)}
    B = 6144
    actor = ma.GLEET_FusedActor(ma.load_reference_weights(ma.GLEET_Actor(), w(G2, 'gleet/')).to(dev).eval())
    x = torch.rand(B, NP, 27, device=dev); nz = torch.randn(B, NP, device=dev)
    line('gleet_policy_kernel', timed(lambda: actor(x, noise=nz), True), B * NP, 112, f'{B} swarms x {NP} particles (L2 flushed)')
    net = ma.DE_DDQN_FusedNet(ma.load_reference_weights(ma.DE_DDQN_Net(), w(G1, 'deddqn/')).to(dev).eval())
    x = torch.rand(4608, 99, device=dev)
    line('deddqn_policy_kernel', timed(lambda: net(x), True), 4608, 400, '4608 lanes (L2 flushed; latency bound: 576 CTAs, one wave)')
    act = ma.RLEPSO_FusedActor(ma.load_reference_weights(ma.RLEPSO_Actor(), w(G1, 'rlepso/')).to(dev).eval())
    x = torch.rand(B, 1, device=dev); nz = torch.randn(B, 35, device=dev)
    line('rlepso_policy_kernel', timed(lambda: act(x, noise=nz), True), B, 4 + 140 + 140, f'{B} lanes (L2 flushed)')
    torch.manual_seed(0)
    pol = ma.RL_PSO_FusedPolicy(ma.RL_PSO_PolicyNet(D).to(dev).eval())          # (the shipped weights are dim 10: random init at dim 30)
    x = torch.rand(B, 2 * D, device=dev); nz = torch.randn(B, device=dev)
    line('rl_pso_policy_kernel', timed(lambda: pol(x, noise=nz), True), B, 8 * D + 8, f'{B} lanes (L2 flushed)')
    qt = ma.QLPSO_FusedQTable(ma.QLPSO_QTable(np.random.RandomState(0).rand(4, 4)).to(dev))
    st = torch.randint(0, 4, (B,), generator=g).to(torch.int32).to(dev); u = torch.rand(B, device=dev)
    line('qlpso_policy_kernel', timed(lambda: qt(st, u=u), True), B, 12, f'{B} lanes (L2 flushed)')
