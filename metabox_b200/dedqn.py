"""DEDQN backbone (DE family, ONE individual per environment step, landscape-feature state) for B lanes in lockstep, and its
single-problem drop-in.  Reference: `src/optimizer/dedqn_optimizer.py:104-206` (+ `random_walk_sampling` :79-90, `cal_reward` :93-101,
`operators/mutate.py:5-9,36-39,88-91,178-181`, `operators/crossover.py:6-18`, `operators/boundary_control.py:5-9`); agent
`src/agent/dedqn_agent.py`.  numpy twin: oracle/dedqn.py (pinned against the unmodified reference, tests/golden/dedqn.npz).

Unlike the other backbones this step is COMPOSED: the two heavy pieces are the library's CUDA kernels -- `mbx_eval` (one trial per
lane, then the whole population again: the reference's `__cal_feature` re-evaluates the population, not the walk) and
`mbx_landscape_features` (fdc / rie / acf / nop of the NP-point walk against those costs) -- and the single-row DE arithmetic between them
(donor of the chosen strategy, clip, binomial crossover, `<=` selection, survival counters, cost-curve protocol) is a handful of batched
torch ops on the same stream.  Nothing synchronises; all state stays on the device.  A fused kernel is the obvious next step once this
composition has been measured.

Storage contract as everywhere: population / costs float32, donor arithmetic float64 in the reference's operation
order, trial rounded to float32 before evaluation.  Accounting like the reference: 2 * NP charged evaluations per step (NP for the single
trial, :172; NP for the features, :135).  Deviations at tolerance level: the walk is a float64 cumulative sum reduced modulo 1 once
(the reference reduces after every addition), the reward is a vectorised float64 sum.
"""
from __future__ import annotations

import numpy as np

from . import _lib

F_, CR_ = 0.5, 0.5                     # dedqn_optimizer.py:107-108
N_ACT = 3


class DEDQN_Engine:
    """`feature` [B,4] float32 (fdc, rie, acf, nop; `feature64` keeps the float64 values); action [B] integer in {0,1,2} = rand/1,
    current-to-rand/1, best/2.  `tape` dicts (parity tests) replay recorded draws: init `init_u [B,NP,D]`, every feature call
    `walk_u [B,NP,D]`, a step `r [B,4]`, `jrand [B]`, `cr_u [B,D]`; without a tape the draws come from the engine's own torch generator
    (`last_tape` keeps what was drawn)."""
    NP = 100

    def __init__(self, pack, lane_inst, maxFEs, n_logpoint=50, log_interval=None, seed=0, NP=None, lane_base=0, _kernels=None, device=None):
        import torch
        self.pack = pack
        if _kernels is None:                                          # product path: CUDA kernels behind the C ABI, no CPU fallback
            _lib.lib()
            from . import landscape
            self.device = pack.device
            self.dim, self.lb, self.ub = pack.dim, float(pack.problems[0].lb), float(pack.problems[0].ub)
            self.has_opt = pack.kind == _lib.KIND_BBOB
            self._eval = lambda x, gen: pack.eval_device(self.lane_inst, x, subtract_optimum=self.has_opt, seed=self._eval_seed, gen=gen)
            self._features = lambda sample, cost: landscape.features(sample, cost)[:, :4]
            hh = getattr(pack, 'h_headers', None)
            self._noisy = True if hh is None or 'noise' not in (hh.dtype.names or ()) else bool((hh['noise'] != 0).any())
        else:                                                         # tests inject CPU stand-ins for the two kernels
            self.device = torch.device(device or 'cpu')
            self.dim, self.lb, self.ub, self.has_opt = _kernels['dim'], float(_kernels['lb']), float(_kernels['ub']), bool(_kernels['has_opt'])
            self._eval, self._features = _kernels['eval'], _kernels['features']
            self._noisy = bool(_kernels.get('noisy', True))
        self.lane_base, self.seed = int(lane_base), int(seed)
        self._eval_seed = (self.seed * 0x9E3779B97F4A7C15 + self.lane_base * 0xD1B54A32D192ED03 + 0x2545F491) & (2 ** 63 - 1)
        self.NP = int(NP or DEDQN_Engine.NP)
        self.B = len(lane_inst)
        self.maxFEs, self.n_logpoint = int(maxFEs), int(n_logpoint)
        self.log_interval = int(log_interval or maxFEs // n_logpoint)
        dev, B, NPn, D = self.device, self.B, self.NP, self.dim
        self.lane_inst = (lane_inst.to(device=dev, dtype=torch.int32) if isinstance(lane_inst, torch.Tensor)
                          else torch.as_tensor(np.asarray(lane_inst, dtype=np.int32)).to(dev)).contiguous()
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)
        self.pop, self.cost, self.survival = z(B, NPn, D), z(B, NPn), z(B, NPn, dt=torch.float64)
        self.gbest, self.gbest_cost, self.gbest_row = z(B, D), z(B), z(B, dt=torch.int64)
        self.fes, self.log_index, self.curve_len = z(B, dt=torch.int32), z(B, dt=torch.int32), z(B, dt=torch.int32)
        self.curve = z(B, _lib.CURVE_LEN)
        self.done = z(B, dt=torch.uint8)
        self.feature, self.feature64, self.reward, self.reward64 = z(B, 4), z(B, 4, dt=torch.float64), z(B), z(B, dt=torch.float64)
        self.pointer, self.step_no = 0, 0
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed((self.seed * 1000003 + self.lane_base * 7919 + 17) & (2 ** 62 - 1))
        self.last_tape = {}

    # ---- random draws ----------------------------------------------------------------------------------------------------------
    def _draw(self, tape, key, make):
        import torch
        if tape is not None and key in tape:
            a = tape[key]
            t = torch.as_tensor(np.ascontiguousarray(a) if isinstance(a, np.ndarray) else a).to(self.device)
        else:
            if tape is not None:
                raise KeyError(f'DEDQN tape without {key!r}')
            t = make()
        self.last_tape[key] = t
        return t

    def _rand(self, *shape):
        import torch
        return torch.rand(*shape, generator=self.gen, device=self.device, dtype=torch.float64)

    # ---- __cal_feature (:123-136) ----------------------------------------------------------------------------------------------
    def _cal_feature(self, tape, gen):
        import torch
        B, NPn, D = self.B, self.NP, self.dim
        u = self._draw(tape, 'walk_u', lambda: self._rand(B, NPn, D)).to(torch.float64)
        walk = torch.remainder(torch.cumsum(u, dim=1), 1.0)                                   # start point, then (point + move) % 1
        x = self.pop.to(torch.float64)
        pmin, pmax = x.min(dim=1, keepdim=True).values, x.max(dim=1, keepdim=True).values
        sample = pmin + (pmax - pmin) * walk
        # the POPULATION's costs, not the walk's (reference quirk).  A noise-free evaluation of the stored population reproduces `self.cost`
        # bit for bit (same kernel, same float32 inputs), so only noisy suites pay for the NP re-evaluations
        c = (self._eval(self.pop, gen) if self._noisy else self.cost).to(torch.float64)
        self.last_sample, self.last_cost = sample, c                                          # (read by the parity tests)
        f = self._features(sample.contiguous(), c.contiguous())
        alive = (self.done == 0)[:, None]
        self.feature64.copy_(torch.where(alive, f, self.feature64))
        self.feature.copy_(self.feature64.to(torch.float32))
        return self.feature

    # ---- init_population (:138-152) --------------------------------------------------------------------------------------------
    def init_population(self, tape=None):
        import torch
        B, NPn, D = self.B, self.NP, self.dim
        self.last_tape = {}
        u = self._draw(tape, 'init_u', lambda: self._rand(B, NPn, D)).to(torch.float64)
        self.pop.copy_((u * (self.ub - self.lb) + self.lb).to(torch.float32))
        self.cost.copy_(self._eval(self.pop, 0))
        self.survival.fill_(1.0)
        self.gbest_row.copy_(torch.argmin(self.cost, dim=1))
        self.gbest.copy_(torch.gather(self.pop, 1, self.gbest_row[:, None, None].expand(B, 1, D))[:, 0])
        self.gbest_cost.copy_(self.cost.min(dim=1).values)
        self.fes.fill_(NPn); self.log_index.fill_(1); self.curve_len.fill_(1)
        self.curve.zero_(); self.curve[:, 0] = self.gbest_cost
        self.done.zero_(); self.reward.zero_(); self.reward64.zero_()
        self.pointer, self.step_no = 0, 0
        self._cal_feature(tape, 1)
        self.fes += NPn
        return self.feature

    # ---- update (:154-206) -----------------------------------------------------------------------------------------------------
    def update(self, action, tape=None, check_action=False):
        """action [B] integer device tensor -> (feature [B,4], reward [B], done [B] uint8), all on the device, no synchronisation
        (`check_action=True` makes the reference's `ValueError('action error')` test, which costs a host sync)."""
        import torch
        B, NPn, D, p = self.B, self.NP, self.dim, self.pointer
        if action.shape != (B,):
            raise ValueError(f'action error: expected shape ({B},), got {tuple(action.shape)}')
        a = action.to(device=self.device, dtype=torch.int64)
        if check_action and bool(((a < 0) | (a >= N_ACT)).any()):
            raise ValueError(f'action error: {action.tolist()}')
        self.last_tape = {}
        self.step_no += 1
        alive = self.done == 0
        # indices != pointer (mutate.py:5-9 rejects whole draws containing the pointer == independent uniforms over the other rows)
        r = self._draw(tape, 'r', lambda: (lambda k: k + (k >= p).to(k.dtype))(
            torch.randint(0, NPn - 1, (B, 4), generator=self.gen, device=self.device))).to(torch.int64)
        jrand = self._draw(tape, 'jrand', lambda: torch.randint(0, D, (B,), generator=self.gen, device=self.device)).to(torch.int64)
        cr_u = self._draw(tape, 'cr_u', lambda: self._rand(B, D)).to(torch.float64)
        x = self.pop.to(torch.float64)
        row = lambda k: torch.gather(x, 1, r[:, k, None, None].expand(B, 1, D))[:, 0]
        x0, x1, x2, x3, xp = row(0), row(1), row(2), row(3), x[:, p]
        best = torch.where((self.gbest_row >= 0)[:, None],
                           torch.gather(x, 1, self.gbest_row.clamp(min=0)[:, None, None].expand(B, 1, D))[:, 0], self.gbest.to(torch.float64))
        v = torch.where((a == 0)[:, None], x0 + F_ * (x1 - x2),                                # rand/1            (mutate.py:36-39)
                        torch.where((a == 1)[:, None], xp + F_ * (x0 - xp + x1 - x2),          # current-to-rand/1 (:178-181)
                                    best + F_ * (x0 - x1 + x2 - x3)))                          # best/2            (:88-91)
        v = torch.clamp(v, self.lb, self.ub)
        take = cr_u < CR_
        take.scatter_(1, jrand[:, None], True)                                                # crossover.py:14-15
        u32 = torch.where(take, v, xp).to(torch.float32)
        u_cost = self._eval(u32[:, None, :].contiguous(), 3 * self.step_no - 1)[:, 0]
        self.fes += torch.where(alive, NPn, 0).to(torch.int32)
        better = alive & (u_cost <= self.cost[:, p])
        self.pop[:, p] = torch.where(better[:, None], u32, self.pop[:, p])
        self.cost[:, p] = torch.where(better, u_cost, self.cost[:, p])
        sp = self.survival[:, p]
        self.survival[:, p] = torch.where(better, torch.ones_like(sp), torch.where(alive, sp + 1, sp))
        gb = better & (u_cost < self.gbest_cost)
        self.gbest.copy_(torch.where(gb[:, None], u32, self.gbest))
        self.gbest_row.copy_(torch.where(gb, torch.full_like(self.gbest_row, -1), self.gbest_row))
        self.gbest_cost.copy_(torch.where(gb, u_cost, self.gbest_cost))
        self._cal_feature(tape, 3 * self.step_no)
        self.fes += torch.where(alive, NPn, 0).to(torch.int32)
        # cost-curve protocol (:183-185, :195-199)
        idx = torch.arange(_lib.CURVE_LEN, device=self.device)[None, :]
        log = alive & (self.fes >= self.log_index * self.log_interval)
        self.log_index += log.to(torch.int32)
        self.curve.copy_(torch.where(log[:, None] & (idx == self.curve_len[:, None]), self.gbest_cost[:, None], self.curve))
        self.curve_len.copy_(torch.clamp(self.curve_len + log.to(torch.int32), max=_lib.CURVE_LEN))
        sv = self.survival
        rew = ((1.0 / sv).sum(dim=1) - 1.0 / sv[:, p] + (sv[:, p] == 1).to(torch.float64)) / NPn          # cal_reward (:93-101)
        self.reward64.copy_(torch.where(alive, rew, torch.zeros_like(rew)))                                 # finished lanes report reward 0
        self.reward.copy_(self.reward64.to(torch.float32))
        fin = self.fes >= self.maxFEs
        if self.has_opt:
            fin = fin | (self.cost.min(dim=1).values.to(torch.float64) <= 1e-8)
        fin = fin & alive
        full = self.curve_len >= self.n_logpoint + 1
        pos = torch.where(full, self.curve_len - 1, self.curve_len)
        self.curve.copy_(torch.where(fin[:, None] & (idx == pos[:, None]), self.gbest_cost[:, None], self.curve))
        self.curve_len.copy_(torch.where(fin & ~full, self.curve_len + 1, self.curve_len))
        self.done.copy_(torch.where(fin, torch.ones_like(self.done), self.done))
        self.pointer = (p + 1) % NPn
        return self.feature, self.reward, self.done

    # ---- results / test access ---------------------------------------------------------------------------------------------------
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

    def read_state(self, b):
        keys = ('pop', 'cost', 'survival', 'gbest', 'gbest_cost', 'gbest_row', 'fes', 'log_index', 'curve_len', 'curve', 'done', 'feature64')
        return {k: getattr(self, k)[b].cpu().numpy().copy() for k in keys}

    def load_state(self, b, **arrays):
        import torch
        for k, a in arrays.items():
            t = getattr(self, k)
            t[b] = torch.as_tensor(np.asarray(a)).to(device=self.device, dtype=t.dtype)


class DEDQN_Optimizer:
    """Drop-in for `DEDQN_Optimizer` (dedqn_optimizer.py:104-206): one problem, numpy in / numpy out, `fes`, `cost`, `log_index`,
    `log_interval` with the reference's meaning; `update` raises `ValueError('action error: ...')` for an action outside {0, 1, 2}."""

    def __init__(self, config):
        config.NP, config.F, config.Cr = DEDQN_Engine.NP, F_, CR_                             # :106-108 (the reference mutates config too)
        config.rwsteps = config.NP
        self.__config = config
        self.log_interval = config.log_interval
        self.fes = self.cost = self.log_index = None
        self._eng = None

    def _sync(self):
        e = self._eng
        self.fes, self.log_index = int(e.fes[0]), int(e.log_index[0])
        self.cost = [float(c) for c in e.cost_curves(pad=False)[0]]

    def init_population(self, problem):
        from .problem import InstancePack
        cfg = self.__config
        pack = problem._pack if getattr(problem, '_pack', None) is not None else InstancePack([problem])      # (sets problem._index)
        self._eng = DEDQN_Engine(pack, [problem._index], cfg.maxFEs, cfg.n_logpoint, self.log_interval, seed=int(np.random.randint(0, 2 ** 31 - 1)))
        self._eng.init_population()
        self._sync()
        return self._eng.feature64[0].cpu().numpy()

    def update(self, action, problem):
        import torch
        if action not in (0, 1, 2):
            raise ValueError(f'action error: {action}')                                       # :165
        e = self._eng
        e.update(torch.tensor([int(action)], device=e.device))
        self._sync()
        return e.feature64[0].cpu().numpy(), float(e.reward64[0]), bool(e.done[0])
