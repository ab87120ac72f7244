#!/usr/bin/env python
"""Throughput of the sharded TRAINING loops (BASELINE config C4 is an LDE train rollout) on the lane machinery: candidate
evaluations per second INCLUDING the policy-gradient update, per trainer.  One process per GPU; under torchrun the problems of a learn step
are sharded over the ranks and the gradients all-reduced once per update.
    python scripts/train_timing.py                                           # one GPU
    python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 scripts/train_timing.py"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from metabox_b200 import problem as mp, trainer as mt   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--dim', type=int, default=30)
ap.add_argument('--steps', type=int, default=3)
ap.add_argument('--only', default='all', choices=['all', 'lde', 'rlepso', 'gleet'])
a = ap.parse_args()
rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
cfg = argparse.Namespace(dim=a.dim, maxFEs=2000 * a.dim, n_logpoint=50, log_interval=2000 * a.dim // 50)


def timed(fn, steps):
    fn()                                                   # warm-up (lazy cuBLAS / kernel attribute initialisation)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    evals = sum(fn() for _ in range(steps))
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1), float(evals)], device='cuda', dtype=torch.float64)
    if world > 1:
        mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        t = torch.stack([mx[0], sm[1]])
    return float(t[1]) / (float(t[0]) / 1000.0), float(t[0]) / steps


if a.only in ('all', 'lde'):                               # C4: bbob-noisy, NP 100, 20 trajectories x 50 steps per problem per learn step
    tr, te = mp.BBOB_Dataset.get_datasets('bbob-noisy', a.dim, 5)
    probs = tr.data + te.data
    T = mt.LDE_Trainer(cfg, probs, NP=100, seed=1)
    ids = list(range(len(probs))) * 2                      # 60 problems per learn step -> 1200 lanes over all ranks
    per_step = len(ids) * mt.TRAJECTORY_NUM * (100 + mt.TRAJECTORY_LENGTH * 100)

    def lde_step():
        T.train_step(ids)
        return per_step / world                            # every rank reports its share
    v, ms = timed(lde_step, a.steps)
    if rank == 0:
        print(f'LDE_Trainer    dim {a.dim} NP 100: {len(ids)} problems x 20 trajectories x 50 steps per learn step, {world} GPU(s): '
              f'{v / 1e6:8.1f} M evals/s incl. the REINFORCE update, {ms:8.1f} ms per learn step')
tr, te = mp.BBOB_Dataset.get_datasets('bbob', a.dim, 5)
probs = tr.data + te.data
ids = list(np.repeat(np.arange(len(probs)), 32))           # 768 lanes (episodes) over all ranks
for name, cls in (('rlepso', mt.RLEPSO_Trainer), ('gleet', mt.GLEET_Trainer)):
    if a.only not in ('all', name):
        continue
    T = cls(cfg, probs, seed=1)

    def episode():
        info = T.train_episodes(ids)
        return float(cfg.maxFEs) * len(info['gbest'])      # upper bound: lanes that converge early charge fewer evaluations
    v, ms = timed(episode, max(1, a.steps - 1))
    if rank == 0:
        print(f'{cls.__name__:14s} dim {a.dim} NP 100: {len(ids)} complete episodes per call (segments of 10 generations x 3 PPO epochs), {world} GPU(s): '
              f'<= {v / 1e6:8.1f} M evals/s incl. the PPO updates, {ms:8.1f} ms per call')
if world > 1:
    dist.destroy_process_group()
