"""Batched rollout driver: the B200 counterpart of the nested `for problem / for run / while not done`
loops of `Tester.test` / `rollout` (`src/tester.py:190-219, 314-346`).

A *lane* is one (instance, run) rollout.  Lanes are independent (fresh seed, `problem.reset()`,
`init_population` per run in the reference), so a sweep is: build the lane table -> shard it over ranks
-> every rank advances its lanes in lockstep, one fused step kernel + one policy forward per generation,
no host synchronisation inside the loop except a done-poll every `poll` generations -> ONE all-gather of
the result rows {curve[51], fes, return} at the end (NCCL over NVLink/NVSwitch).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib
from .optimizer import RLEPSO_Engine


def lane_table(n_instances, n_runs):
    """(instance id, run id) per lane, instance-major: lane = inst * n_runs + run."""
    inst = np.repeat(np.arange(n_instances, dtype=np.int32), n_runs)
    run = np.tile(np.arange(n_runs, dtype=np.int32), n_instances)
    return inst, run


def shard_range(n_lanes, rank, world):
    """Contiguous block of lanes for `rank` (sizes differ by at most 1)."""
    q, r = divmod(n_lanes, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


class RLEPSO_Sweep:
    """All lanes of one rank through complete RLEPSO rollouts with the policy on the same stream."""

    def __init__(self, pack, lane_inst, actor, maxFEs, n_logpoint=50, seed=0, lane_base=0, poll=64, time_kernels=False):
        self.engine = RLEPSO_Engine(pack, lane_inst, maxFEs, n_logpoint, seed=seed, lane_base=lane_base)
        self.actor = actor
        self.poll = poll
        self.max_gens = int(math.ceil(maxFEs / self.engine.NP)) + 1
        self.time_kernels = time_kernels
        self.kernel_events = []
        self.launches = 0
        self.ret = torch.zeros(self.engine.B, dtype=torch.float32, device=self.engine.device)

    @torch.no_grad()
    def run(self):
        """-> dict(curve [B,51] f32, curve_len, fes i32, ret f32) device tensors (no host copy)."""
        eng = self.engine
        self.ret.zero_()
        self.kernel_events = []
        feat = eng.init_population()
        self.launches = 1
        gens = 0
        while gens < self.max_gens:
            for _ in range(self.poll):
                alive = (eng.done == 0).to(torch.float32)
                action = self.actor(feat)
                if self.time_kernels:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    feat, reward, done = eng.update(action)
                    e1.record()
                    self.kernel_events.append((e0, e1))
                else:
                    feat, reward, done = eng.update(action)
                self.ret += reward * alive
                self.launches += 1
                gens += 1
                if gens >= self.max_gens:
                    break
            if bool(torch.all(eng.done != 0)):          # the only host sync inside the sweep
                break
        # cost curves padded to 51 like tester.py:204-205
        idx = torch.arange(_lib.CURVE_LEN, device=eng.device)[None, :]
        last = torch.gather(eng.curve, 1, (eng.curve_len.long() - 1).clamp(min=0)[:, None])
        curve = torch.where(idx < eng.curve_len[:, None], eng.curve, last)
        return dict(curve=curve, curve_len=eng.curve_len, fes=eng.fes, ret=self.ret, gens=gens)

    def kernel_ms(self):
        """Sum / count of the step-kernel durations recorded by the last run (needs time_kernels=True)."""
        if not self.kernel_events:
            return 0.0, 0
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in self.kernel_events), len(self.kernel_events)


def gather_results(res, world):
    """ONE all-gather of the packed result rows [n_local, 51+2] over the default process group."""
    import torch.distributed as dist
    row = torch.cat([res['curve'], res['fes'].to(torch.float32)[:, None], res['ret'][:, None]], dim=1).contiguous()
    if world == 1:
        return row
    n = torch.tensor([row.shape[0]], device=row.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    mx = int(max(int(s) for s in sizes))
    pad = torch.zeros(mx, row.shape[1], device=row.device, dtype=row.dtype)
    pad[:row.shape[0]] = row
    out = torch.empty(world * mx, row.shape[1], device=row.device, dtype=row.dtype)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * mx: r * mx + int(sizes[r])] for r in range(world)], dim=0)


def results_to_reference_dict(rows, problems, n_runs, agent_name):
    """rows [n_lanes, 53] (host) -> the reference's `test_results` layout (tester.py:123-127):
    {'cost': {problem: {agent: [runs x 51]}}, 'fes': {problem: {agent: [runs]}}}, plus 'return'."""
    rows = np.asarray(rows)
    out = {'cost': {}, 'fes': {}, 'return': {}}
    for k, p in enumerate(problems):
        blk = rows[k * n_runs:(k + 1) * n_runs]
        out['cost'][str(p)] = {agent_name: [list(map(float, r[:51])) for r in blk]}
        out['fes'][str(p)] = {agent_name: [int(r[51]) for r in blk]}
        out['return'][str(p)] = {agent_name: [float(r[52]) for r in blk]}
    return out
