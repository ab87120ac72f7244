"""RL_HPSDE backbone (DE family: whole population per step, success-history memories of F / Cr, Cauchy / Levy step lengths, linear
population-size reduction, landscape state from a progressive random walk) for B lanes in lockstep, and its single-problem drop-in.
Reference: `src/optimizer/rl_hpsde_optimizer.py:7-287` (+ `operators/mutate.py:12-32,148-157,184-188`, `operators/crossover.py:6-18`);
agent `src/agent/rl_hpsde_agent.py` (tabular Q-learning over the 4 states x 4 actions).  numpy twin: oracle/rl_hpsde.py (pinned against the
unmodified reference, tests/golden/rl_hpsde.npz).

Like DEDQN (dedqn.py) this step is COMPOSED of library kernels and batched torch ops on one stream, nothing synchronises:
  `mbx_eval`                 all NP trials of every lane, then the 201 walk points
  `mbx_progressive_walk`     the 200 sequential, reflecting walk moves (one thread per lane and dimension)
  `mbx_landscape_features`   dynamic fitness-distance correlation and ruggedness entropy of the walk (columns dfdc_r, drie_r)
  torch                      memory look-ups, donor vectors of the two strategies, binomial crossover, strict selection, weighted Lehmer
                             means, stable sort by cost, truncation to the scheduled population size, cost-curve protocol.
All lanes share the population-size schedule: every generation charges NP + 201 evaluations to every running lane, so NP(t) is one
host-side integer and the [B, NP, D] tensors are physically truncated; a lane that finishes early (cost <= 1e-8) freezes.

Storage contract as everywhere: population / costs float32, donor arithmetic float64 in the reference's operation order, trials and walk
points rounded to float32 before evaluation.  Tied costs are sorted in index order (stable); the reference's quicksort order is unspecified.
"""
from __future__ import annotations

import math

import numpy as np

from . import _lib

RW_STEPS, STEP_SIZE, NP_MIN = 200, 10.0, 4                      # rl_hpsde_optimizer.py:108-111
N_ACT = 4


class RL_HPSDE_Engine:
    """`feature` [B] int32 in 0..3 (= [DFDC r > 0.15] + 2 [DRIE entropy >= 0.5]); action [B] integer in 0..3 (Cauchy / Levy step lengths x
    current-to-rand/1 / current-to-best/1).  `tape` dicts (parity tests) replay recorded draws: init `init_u [B,NP,D]`; every state
    computation `zone_u, r_u [B,D]`, `rD [B]`, `step_u [B,200,D]`; a generation `ind_r, cr_raw, f_raw [B,NP]`, `r [B,NP,3]`,
    `jrand [B,NP]`, `cr_u [B,NP,D]`.  Without a tape the draws come from the engine's own torch generator (`last_tape` keeps them)."""

    def __init__(self, pack, lane_inst, maxFEs, n_logpoint=50, log_interval=None, seed=0, lane_base=0, _kernels=None, device=None):
        import torch
        self.pack = pack
        if _kernels is None:                                          # product path: CUDA kernels behind the C ABI, no CPU fallback
            _lib.lib()
            self.device = pack.device
            self.dim, self.lb, self.ub = pack.dim, float(pack.problems[0].lb), float(pack.problems[0].ub)
            self.has_opt = pack.kind == _lib.KIND_BBOB
            self._eval = lambda x, gen: pack.eval_device(self.lane_inst, x, subtract_optimum=self.has_opt, seed=self._eval_seed, gen=gen)
            self._features = self._features_cuda
            self._walk = self._walk_cuda
        else:                                                         # tests inject CPU stand-ins for the three kernels
            self.device = torch.device(device or 'cpu')
            self.dim, self.lb, self.ub, self.has_opt = _kernels['dim'], float(_kernels['lb']), float(_kernels['ub']), bool(_kernels['has_opt'])
            self._eval, self._features, self._walk = _kernels['eval'], _kernels['features'], _kernels['walk']
        self.lane_base, self.seed = int(lane_base), int(seed)
        self._eval_seed = (self.seed * 0x9E3779B97F4A7C15 + self.lane_base * 0xD1B54A32D192ED03 + 0x51ED27) & (2 ** 63 - 1)
        self.B = len(lane_inst)
        self.NP_max, self.H = 18 * self.dim, int(self.dim * 0.5)
        self.maxFEs, self.n_logpoint = int(maxFEs), int(n_logpoint)
        self.log_interval = int(log_interval or maxFEs // n_logpoint)
        dev, B = self.device, self.B
        self.lane_inst = (lane_inst.to(device=dev, dtype=torch.int32) if isinstance(lane_inst, torch.Tensor)
                          else torch.as_tensor(np.asarray(lane_inst, dtype=np.int32)).to(dev)).contiguous()
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)
        self.group, self.cost = z(B, self.NP_max, self.dim), z(B, self.NP_max)
        self.MF, self.MCr, self.k = z(B, self.H, dt=torch.float64), z(B, self.H, dt=torch.float64), z(B, dt=torch.int64)
        self.gbest = z(B)
        self.fes, self.log_index, self.curve_len = z(B, dt=torch.int32), z(B, dt=torch.int32), z(B, dt=torch.int32)
        self.curve = z(B, _lib.CURVE_LEN)
        self.done = z(B, dt=torch.uint8)
        self.feature, self.reward, self.reward64 = z(B, dt=torch.int32), z(B), z(B, dt=torch.float64)
        self.dfdc, self.drie = z(B, dt=torch.float64), z(B, dt=torch.float64)
        self.fes_running, self.step_no = 0, 0                          # evaluations charged to every lane that is still running
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed((self.seed * 1000003 + self.lane_base * 7919 + 29) & (2 ** 62 - 1))
        self.last_tape = {}

    NP = property(lambda self: self.cost.shape[1])

    # ---- kernels (product path) ------------------------------------------------------------------------------------------------
    def _features_cuda(self, sample, cost, ref):
        from . import landscape
        return landscape.features(sample, cost, ref)[:, 4:6]

    def _walk_cuda(self, zone_u, r_u, rD, step_u):
        import torch
        B, D = zone_u.shape
        out = torch.empty(B, RW_STEPS + 1, D, dtype=torch.float64, device=self.device)
        out32 = torch.empty(B, RW_STEPS + 1, D, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            rc = _lib.lib().mbx_progressive_walk(_lib.ptr(zone_u.contiguous()), _lib.ptr(r_u.contiguous()), _lib.ptr(rD.to(torch.int32).contiguous()),
                                                 _lib.ptr(step_u.contiguous()), self.lb, self.ub, STEP_SIZE, B, D, RW_STEPS, _lib.ptr(out),
                                                 _lib.ptr(out32), _lib.stream_ptr())
        _lib.check(rc, 'mbx_progressive_walk')
        return out32

    # ---- random draws ----------------------------------------------------------------------------------------------------------
    def _draw(self, tape, key, make):
        import torch
        if tape is not None:
            if key not in tape:
                raise KeyError(f'RL_HPSDE tape without {key!r}')
            a = tape[key]
            t = torch.as_tensor(np.ascontiguousarray(a) if isinstance(a, np.ndarray) else a).to(self.device)
        else:
            t = make()
        self.last_tape[key] = t
        return t

    def _rand(self, *shape):
        import torch
        return torch.rand(*shape, generator=self.gen, device=self.device, dtype=torch.float64)

    def _randn(self, *shape):
        import torch
        return torch.randn(*shape, generator=self.gen, device=self.device, dtype=torch.float64)

    def _distinct(self, N):
        """[B,N,3] indices, mutually distinct and != the row index, uniform over all such triples (what the rejection loop of
        `generate_random_int`, mutate.py:12-32, converges to): draw from the shrinking set of allowed values, skipping the excluded ones in
        ascending order."""
        import torch
        B = self.B
        if N < 4:       # only past maxFEs (the generation that crosses the budget shrinks below 4, like the reference's): every lane has
            return torch.zeros(B, N, 3, dtype=torch.int64, device=self.device)     # finished, a sweep's surplus step must not draw
        i = torch.arange(N, device=self.device)[None, :].expand(B, N)
        pick = lambda n: torch.randint(0, n, (B, N), generator=self.gen, device=self.device)
        r0 = pick(N - 1); r0 = r0 + (r0 >= i).to(r0.dtype)
        ex = torch.sort(torch.stack([i, r0], dim=-1), dim=-1).values
        r1 = pick(N - 2); r1 = r1 + (r1 >= ex[..., 0]).to(r1.dtype); r1 = r1 + (r1 >= ex[..., 1]).to(r1.dtype)
        ex = torch.sort(torch.stack([i, r0, r1], dim=-1), dim=-1).values
        r2 = pick(N - 3)
        for c in range(3):
            r2 = r2 + (r2 >= ex[..., c]).to(r2.dtype)
        return torch.stack([r0, r1, r2], dim=-1)

    # ---- __get_state (:218-229) ------------------------------------------------------------------------------------------------
    def _get_state(self, tape, gen, alive=None):
        import torch
        B, D = self.B, self.dim
        zone_u = self._draw(tape, 'zone_u', lambda: self._rand(B, D)).to(torch.float64)
        r_u = self._draw(tape, 'r_u', lambda: self._rand(B, D)).to(torch.float64)
        rD = self._draw(tape, 'rD', lambda: torch.randint(0, D, (B,), generator=self.gen, device=self.device))
        step_u = self._draw(tape, 'step_u', lambda: self._rand(B, RW_STEPS, D)).to(torch.float64)
        sample32 = self._walk(zone_u, r_u, rD, step_u)                                       # [B,201,D] f32: the points as evaluated
        c = self._eval(sample32, gen).to(torch.float64)
        best = self.group[:, 0].to(torch.float64)                                            # sorted population: row 0 is gbest_solution
        self.last_sample, self.last_cost = sample32, c                                       # (read by the parity tests)
        f = self._features(sample32.to(torch.float64).contiguous(), c.contiguous(), best.contiguous())
        alive = (self.done == 0) if alive is None else alive                                 # lanes running when the generation started
        self.dfdc.copy_(torch.where(alive, f[:, 0], self.dfdc)); self.drie.copy_(torch.where(alive, f[:, 1], self.drie))
        st = (self.dfdc > 0.15).to(torch.int32) + 2 * (self.drie >= 0.5).to(torch.int32)
        self.feature.copy_(st)
        return self.feature

    def _sort(self):
        import torch
        self.cost, ind = torch.sort(self.cost, dim=1, stable=True)
        self.group = torch.gather(self.group, 1, ind[:, :, None].expand(-1, -1, self.dim))
        self.gbest.copy_(self.cost[:, 0])

    # ---- init_population (:126-134) --------------------------------------------------------------------------------------------
    def init_population(self, tape=None):
        import torch
        B, N, D = self.B, self.NP_max, self.dim
        self.last_tape = {}
        u = self._draw(tape, 'init_u', lambda: self._rand(B, N, D)).to(torch.float64)
        self.group = (u * (self.ub - self.lb) + self.lb).to(torch.float32)
        self.cost = self._eval(self.group, 0)
        self.MF.fill_(0.5); self.MCr.fill_(0.5); self.k.zero_()
        self._sort()
        self.fes.fill_(N); self.log_index.fill_(1); self.curve_len.fill_(1)
        self.curve.zero_(); self.curve[:, 0] = self.gbest
        self.done.zero_(); self.reward.zero_(); self.reward64.zero_()
        self.fes_running, self.step_no = N, 0
        self._get_state(tape, 1)
        self.fes += RW_STEPS + 1
        self.fes_running += RW_STEPS + 1
        return self.feature

    # ---- update (:231-287) -----------------------------------------------------------------------------------------------------
    def update(self, action, tape=None, check_action=False):
        """action [B] integer device tensor -> (feature [B] int32, reward [B] f32, done [B] uint8), all on the device, no synchronisation
        (`check_action=True` makes the reference's `ValueError('action error')` test, which costs a host sync)."""
        import torch
        B, N, D = self.B, self.NP, self.dim
        if action.shape != (B,):
            raise ValueError(f'action error: expected shape ({B},), got {tuple(action.shape)}')
        a = action.to(device=self.device, dtype=torch.int64)
        if check_action and bool(((a < 0) | (a >= N_ACT)).any()):
            raise ValueError(f'action error: {action.tolist()}')
        self.last_tape = {}
        self.step_no += 1
        alive = self.done == 0
        levy = (a >= 2)[:, None]
        to_best = ((a == 1) | (a == 3))[:, None, None]
        # step lengths and crossover rates from the memories (:52-68)
        ind_r = self._draw(tape, 'ind_r', lambda: torch.randint(0, self.H, (B, N), generator=self.gen, device=self.device)).to(torch.int64)
        locs, mcr = torch.gather(self.MF, 1, ind_r), torch.gather(self.MCr, 1, ind_r)
        cr_raw = self._draw(tape, 'cr_raw', lambda: mcr + 0.1 * self._randn(B, N)).to(torch.float64)
        f_raw = self._draw(tape, 'f_raw', lambda: torch.where(levy, locs + 0.1 / self._randn(B, N) ** 2,                # Levy = 1 / Z^2
                                                              locs + 0.1 * torch.tan(math.pi * (self._rand(B, N) - 0.5)))).to(torch.float64)
        Cr = torch.clamp(cr_raw, 0.0, 1.0)
        F = torch.where(f_raw < 0, 2 * locs - f_raw, f_raw)
        F = torch.clamp(F, max=1.0)[:, :, None]
        r = self._draw(tape, 'r', lambda: self._distinct(N)).to(torch.int64)
        jrand = self._draw(tape, 'jrand', lambda: torch.randint(0, D, (B, N), generator=self.gen, device=self.device)).to(torch.int64)
        cr_u = self._draw(tape, 'cr_u', lambda: self._rand(B, N, D)).to(torch.float64)
        x = self.group.to(torch.float64)
        rows = lambda k: torch.gather(x, 1, r[:, :, k, None].expand(B, N, D))
        x0, x1, x2 = rows(0), rows(1), rows(2)
        best = x[:, :1]                                                                       # sorted: row 0 = gbest_solution
        v = torch.where(to_best, x + F * (best - x + x0 - x1),                                # current-to-best/1 (mutate.py:148-157)
                        x + F * (x0 - x + x1 - x2))                                           # current-to-rand/1 (mutate.py:184-188)
        v = torch.clamp(v, self.lb, self.ub)
        take = cr_u < Cr[:, :, None]
        take.scatter_(2, jrand[:, :, None], True)                                             # crossover.py:14-15
        u32 = torch.where(take, v, x).to(torch.float32)
        ncost = self._eval(u32.contiguous(), 2 * self.step_no)
        self.fes += torch.where(alive, N, 0).to(torch.int32)
        better = alive[:, None] & (ncost < self.cost)
        # success-history memories (:70-87): weighted Lehmer means of the successful F / Cr, weights = cost improvements
        df = torch.where(better, self.cost.to(torch.float64) - ncost.to(torch.float64), torch.zeros((), dtype=torch.float64, device=self.device))
        w = df / df.sum(dim=1, keepdim=True).clamp(min=1e-300)
        lehmer = lambda s: torch.where((w * s).sum(1) > 0.000001, (w * s * s).sum(1) / (w * s).sum(1).clamp(min=1e-300), torch.full_like(s[:, 0], 0.5))
        any_ok = better.any(dim=1)
        slot = torch.nn.functional.one_hot(self.k, self.H).to(torch.bool)
        newF = torch.where(any_ok, lehmer(F[:, :, 0]), torch.full_like(self.gbest, 0.5, dtype=torch.float64))
        newCr = torch.where(any_ok, lehmer(Cr), torch.full_like(self.gbest, 0.5, dtype=torch.float64))
        upd = slot & alive[:, None]
        self.MF.copy_(torch.where(upd, newF[:, None], self.MF)); self.MCr.copy_(torch.where(upd, newCr[:, None], self.MCr))
        self.k.copy_(torch.where(alive & any_ok, (self.k + 1) % self.H, self.k))
        self.group = torch.where(better[:, :, None], u32, self.group)
        self.cost = torch.where(better, ncost, self.cost)
        rew = better.sum(dim=1).to(torch.float64) / N
        # LPSR (:90-99): sort, then truncate to the scheduled size (one schedule for all running lanes)
        self._sort()
        self.fes_running += N
        N_new = max(int(self.NP_max + round(NP_MIN - self.NP_max) * self.fes_running / self.maxFEs), 1)
        if N_new < N:
            self.group, self.cost = self.group[:, :N_new].contiguous(), self.cost[:, :N_new].contiguous()
        # cost-curve protocol (:262-275)
        idx = torch.arange(_lib.CURVE_LEN, device=self.device)[None, :]
        log = alive & (self.fes >= self.log_index * self.log_interval)
        self.log_index += log.to(torch.int32)
        self.curve.copy_(torch.where(log[:, None] & (idx == self.curve_len[:, None]), self.gbest[:, None], self.curve))
        self.curve_len.copy_(torch.clamp(self.curve_len + log.to(torch.int32), max=_lib.CURVE_LEN))
        self.reward64.copy_(torch.where(alive, rew, torch.zeros_like(rew)))
        self.reward.copy_(self.reward64.to(torch.float32))
        fin = self.fes >= self.maxFEs
        if self.has_opt:
            fin = fin | (self.gbest.to(torch.float64) <= 1e-8)
        fin = fin & alive
        full = self.curve_len >= self.n_logpoint + 1
        pos = torch.where(full, self.curve_len - 1, self.curve_len)
        self.curve.copy_(torch.where(fin[:, None] & (idx == pos[:, None]), self.gbest[:, None], self.curve))
        self.curve_len.copy_(torch.where(fin & ~full, self.curve_len + 1, self.curve_len))
        self.done.copy_(torch.where(fin, torch.ones_like(self.done), self.done))
        # the state after the generation (the reference computes it even when the run has just ended)
        self._get_state(tape, 2 * self.step_no + 1, alive)
        self.fes += torch.where(alive, RW_STEPS + 1, 0).to(torch.int32)
        self.fes_running += RW_STEPS + 1
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
        keys = ('group', 'cost', 'MF', 'MCr', 'k', 'gbest', 'fes', 'log_index', 'curve_len', 'curve', 'done', 'feature', 'dfdc', 'drie')
        return {k: getattr(self, k)[b].cpu().numpy().copy() for k in keys}


class RL_HPSDE_Optimizer:
    """Drop-in for `RL_HPSDE_Optimizer` (rl_hpsde_optimizer.py:102-287): one problem, numpy in / python out; `update` raises
    `ValueError('action error: ...')` for an action outside {0, 1, 2, 3}."""

    def __init__(self, config):
        config.F, config.Cr, config.NP_max, config.NP_min, config.Hm = 0.5, 0.5, 18 * config.dim, NP_MIN, 0.5      # :105-109
        config.rw_steps, config.step_size = RW_STEPS, int(STEP_SIZE)
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
        self._eng = RL_HPSDE_Engine(pack, [problem._index], cfg.maxFEs, cfg.n_logpoint, self.log_interval, seed=int(np.random.randint(0, 2 ** 31 - 1)))
        self._eng.init_population()
        self._sync()
        return int(self._eng.feature[0])

    def update(self, action, problem):
        import torch
        if action not in (0, 1, 2, 3):
            raise ValueError(f'action error: {action}')                                       # :248
        e = self._eng
        e.update(torch.tensor([int(action)], device=e.device))
        self._sync()
        return int(e.feature[0]), float(e.reward64[0]), bool(e.done[0])
