"""L2L backbone (a learned point-wise optimizer: the agent's LSTM proposes ONE point per step, 100 steps per episode) for B lanes, its sweep
and its single-problem drop-in.  Reference: `src/optimizer/l2l_optimizer.py:10-53` (sigmoid scaling into the box, evaluate, best-so-far
list, done after 100 evaluations or at cost <= 1e-8) and `src/agent/l2l_agent.py:97-134` (`rollout_episode`).  numpy twin: oracle/l2l.py
(pinned against whole episodes of the unmodified reference, tests/golden/l2l.npz).

The step is one `mbx_eval` launch ([B, 1, D]) plus a handful of torch ops; the policy (`agent.L2L_PolicyNet`, float64 like the reference)
is a written-out LSTM cell with projection on the same stream.  The sweep runs eagerly: `mbx_eval` takes the step index (the key of the
noise stream of the noisy functions) as a host argument, which a captured graph would freeze.
"""
from __future__ import annotations

import numpy as np

from . import _lib

T_MAX = 100


class L2L_Engine:
    """`feature` [B, D+2] float64 = [previous raw action | previous cost | 1] (zeros before the first step); action [B, D] float64 (raw LSTM
    output, scaled into the box by a sigmoid here); `update` -> (feature, y [B] float32 = this step's cost, done [B] uint8).
    `curve` holds every second best-so-far value (what `rollout_episode` returns as 'cost'), `ysum / y0` is its 'return'."""
    NP = 1

    def __init__(self, pack, lane_inst, maxFEs=T_MAX, n_logpoint=50, log_interval=None, seed=0, lane_base=0, _kernels=None, device=None):
        import torch
        self.pack = pack
        if _kernels is None:                                          # product path: the CUDA kernel behind the C ABI, no CPU fallback
            _lib.lib()
            self.device = pack.device
            self.dim, self.lb, self.ub = pack.dim, float(pack.problems[0].lb), float(pack.problems[0].ub)
            self.has_opt = pack.kind == _lib.KIND_BBOB
            self._eval = lambda x, gen: pack.eval_device(self.lane_inst, x, subtract_optimum=self.has_opt, seed=self._eval_seed, gen=gen)
        else:
            self.device = torch.device(device or 'cpu')
            self.dim, self.lb, self.ub, self.has_opt = _kernels['dim'], float(_kernels['lb']), float(_kernels['ub']), bool(_kernels['has_opt'])
            self._eval = _kernels['eval']
        self.lane_base, self.seed = int(lane_base), int(seed)
        self._eval_seed = (self.seed * 0x9E3779B97F4A7C15 + self.lane_base * 0xD1B54A32D192ED03 + 0x77) & (2 ** 63 - 1)
        self.B = len(lane_inst)
        dev, B = self.device, self.B
        self.lane_inst = (lane_inst.to(device=dev, dtype=torch.int32) if isinstance(lane_inst, torch.Tensor)
                          else torch.as_tensor(np.asarray(lane_inst, dtype=np.int32)).to(dev)).contiguous()
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)
        self.feature = z(B, self.dim + 2, dt=torch.float64)
        self.best, self.y0, self.ysum, self.reward = z(B), z(B, dt=torch.float64), z(B, dt=torch.float64), z(B)
        self.fes, self.curve_len = z(B, dt=torch.int32), z(B, dt=torch.int32)
        self._t = 0
        self.curve = z(B, _lib.CURVE_LEN)
        self.done = z(B, dt=torch.uint8)

    def init_population(self):
        self.feature.zero_(); self.best.fill_(float('inf')); self.y0.zero_(); self.ysum.zero_(); self.reward.zero_()
        self.fes.zero_(); self.curve_len.zero_(); self.curve.zero_(); self.done.zero_()
        self._t = 0
        return self.feature

    def update(self, action):
        import torch
        if action.shape != (self.B, self.dim):
            raise ValueError(f'action must be [B, dim] = ({self.B}, {self.dim}), got {tuple(action.shape)}')
        raw = action.to(device=self.device, dtype=torch.float64)
        alive = self.done == 0
        x32 = (self.lb + (self.ub - self.lb) * torch.sigmoid(raw)).to(torch.float32)          # np_scale (:10-13), then the storage contract
        self._t += 1                                                                          # host-side step index: keys the noise stream of noisy
        y = self._eval(x32[:, None, :].contiguous(), self._t)[:, 0]                           # functions (so the sweep runs this engine eagerly)
        first = alive & (self.fes == 0)
        self.best.copy_(torch.where(alive & (y < self.best), y, self.best))
        even = alive & (self.fes % 2 == 0)                                                    # cost[::2] of the best-so-far list
        idx = torch.arange(_lib.CURVE_LEN, device=self.device)[None, :]
        self.curve.copy_(torch.where(even[:, None] & (idx == self.curve_len[:, None]), self.best[:, None], self.curve))
        self.curve_len += even.to(torch.int32)
        self.fes += alive.to(torch.int32)
        y64 = y.to(torch.float64)
        self.y0.copy_(torch.where(first, y64, self.y0))
        self.ysum += torch.where(alive, y64, torch.zeros_like(y64))
        self.reward.copy_(torch.where(alive, y, torch.zeros_like(y)))
        fin = self.fes >= T_MAX
        if self.has_opt:
            fin = fin | (y64 <= 1e-8)
        self.feature.copy_(torch.where(alive[:, None], torch.cat((raw, y64[:, None], torch.ones_like(y64)[:, None]), dim=1), self.feature))
        self.done.copy_(torch.where(fin & alive, torch.ones_like(self.done), self.done))
        return self.feature, self.reward, self.done

    def episode_return(self):
        """`y_sum / init_y` (l2l_agent.py:134)."""
        return self.ysum / self.y0

    def cost_curves(self, pad=True):
        cur = self.curve.cpu().numpy().astype(np.float64)
        ln = self.curve_len.cpu().numpy()
        out = []
        for b in range(self.B):
            c = list(cur[b, :ln[b]])
            if pad:
                while len(c) < _lib.CURVE_LEN:
                    c.append(c[-1])
            out.append(c)
        return out


class L2L_Optimizer:
    """Drop-in for `L2L_Optimizer` (l2l_optimizer.py:15-53): `update(action [dim] numpy, problem) -> (y, 0, is_done)`; `cost` is the
    best-so-far list.  (The reference also accepts a torch tensor with gradients for training on the `-torch` problems; only the numpy
    rollout path exists here.)"""

    def __init__(self, config):
        self.__config = config
        self.cost = None
        self._eng = None

    def init_population(self, problem):
        from .problem import InstancePack
        pack = problem._pack if getattr(problem, '_pack', None) is not None else InstancePack([problem])
        self._eng = L2L_Engine(pack, [problem._index])
        self._eng.init_population()
        self.cost = []

    def update(self, action, problem):
        import torch
        e = self._eng
        _, y, done = e.update(torch.as_tensor(np.asarray(action, dtype=np.float64).reshape(1, -1)).to(e.device))
        self.cost.append(float(e.best[0]))
        return np.float64(float(y[0])), 0, bool(done[0])
