"""NRLPSO backbone (PSO family, ONE particle per environment step: neighbourhood-based learning targets chosen by one of four strategies,
neighbourhood differential mutation for stagnating particles, crowding-aware reward) for B lanes in lockstep, and its single-problem
drop-in.  Reference: `src/optimizer/nrlpso_optimizer.py:7-297`; agent `src/agent/nrlpso_agent.py` (tabular Q-learning over the 4
strategies).  numpy twin: oracle/nrlpso.py (pinned against the unmodified reference incl. its quirks -- `gbest_pos` as a view of a
population row, personal-best costs that are never refreshed, stale neighbourhood snapshots -- tests/golden/nrlpso.npz).

COMPOSED like dedqn.py / rl_hpsde.py: `mbx_eval` evaluates the moved particle and the two neighbourhood mutants of every lane
([B, 1, D] each), everything else -- the pairwise-distance matrices of the crowding term (twice per step), the neighbourhood tables (rebuilt
when the pointer wraps), the masked velocity rule of the four strategies -- is batched torch on the same stream; nothing synchronises.
Every lane evaluates both mutants every step and uses them only where the reference would (stagnation count >= 2); charged evaluations
follow the reference (1 per step, +2 with a mutation).
"""
from __future__ import annotations

import numpy as np

from . import _lib

NP_, K_, N_STATE = 100, 5, 4                                   # nrlpso_optimizer.py:12-13, 20
W_MAX, W_MIN, U_, V_ = 1.0, 0.4, 0.6, 0.33                     # :15-18
_C1 = (2.2, 2.1, 2.0, 1.8)                                     # :146, 159, 172, 186
_C2 = (1.8, 1.8, 2.0, 2.2)


class NRLPSO_Engine:
    """`feature` [B] int32 = the stored strategy of the particle the next action applies to; action [B] integer in 0..3; reward in
    {2, 1, 0, -2}.  `tape` dicts replay recorded draws (init: `init_u [B,NP,D]`, `r_w [B]`, `qstate [B,NP]`; a step: `r1, r2 [B]`, `pick_a,
    pick_b [B]`, `r1v, r2v, mut_p, mut_g [B,D]`); without a tape they come from the engine's own torch generator (`last_tape`)."""
    NP = NP_

    def __init__(self, pack, lane_inst, maxFEs, n_logpoint=50, log_interval=None, seed=0, lane_base=0, _kernels=None, device=None):
        import torch
        self.pack = pack
        if _kernels is None:                                          # product path: the CUDA kernel behind the C ABI, no CPU fallback
            _lib.lib()
            self.device = pack.device
            self.dim, self.lb, self.ub = pack.dim, float(pack.problems[0].lb), float(pack.problems[0].ub)
            self.has_opt = pack.kind == _lib.KIND_BBOB
            self._eval = lambda x, gen: pack.eval_device(self.lane_inst, x, subtract_optimum=self.has_opt, seed=self._eval_seed, gen=gen)
        else:                                                         # tests inject a CPU stand-in for mbx_eval
            self.device = torch.device(device or 'cpu')
            self.dim, self.lb, self.ub, self.has_opt = _kernels['dim'], float(_kernels['lb']), float(_kernels['ub']), bool(_kernels['has_opt'])
            self._eval = _kernels['eval']
        self.lane_base, self.seed = int(lane_base), int(seed)
        self._eval_seed = (self.seed * 0x9E3779B97F4A7C15 + self.lane_base * 0xD1B54A32D192ED03 + 0xA11CE) & (2 ** 63 - 1)
        self.B = len(lane_inst)
        self.maxFEs, self.n_logpoint = int(maxFEs), int(n_logpoint)
        self.log_interval = int(log_interval or maxFEs // n_logpoint)
        dev, B, D = self.device, self.B, self.dim
        self.lane_inst = (lane_inst.to(device=dev, dtype=torch.int32) if isinstance(lane_inst, torch.Tensor)
                          else torch.as_tensor(np.asarray(lane_inst, dtype=np.int32)).to(dev)).contiguous()
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)
        self.pop, self.vel, self.pbest_pos = z(B, NP_, D), z(B, NP_, D), z(B, NP_, D)
        self.cost, self.pbest_cost = z(B, NP_), z(B, NP_)
        self.gbest_pos, self.gbest_cost, self.gbest_row = z(B, D), z(B), z(B, dt=torch.int64)
        self.stag, self.qstate = z(B, NP_, dt=torch.int64), z(B, NP_, dt=torch.int64)
        self.r_w, self.w = z(B, dt=torch.float64), z(B, dt=torch.float64)
        self.pbest_neb_index, self.pbest_neb = z(B, NP_, K_, dt=torch.int64), z(B, NP_, K_, D)
        self.gbest_neb_index, self.gbest_neb = z(B, K_, dt=torch.int64), z(B, K_, D)
        self.fes, self.log_index, self.curve_len = z(B, dt=torch.int32), z(B, dt=torch.int32), z(B, dt=torch.int32)
        self.curve = z(B, _lib.CURVE_LEN)
        self.done = z(B, dt=torch.uint8)
        self.feature, self.reward = z(B, dt=torch.int32), z(B)
        self.pointer, self.step_no = 0, 0
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed((self.seed * 1000003 + self.lane_base * 7919 + 41) & (2 ** 62 - 1))
        self.last_tape = {}

    # ---- helpers ---------------------------------------------------------------------------------------------------------------
    def _draw(self, tape, key, make):
        import torch
        if tape is not None:
            if key not in tape:
                raise KeyError(f'NRLPSO tape without {key!r}')
            a = tape[key]
            t = torch.as_tensor(np.ascontiguousarray(a) if isinstance(a, np.ndarray) else a).to(self.device)
        else:
            t = make()
        self.last_tape[key] = t
        return t

    def _rand(self, *shape):
        import torch
        return torch.rand(*shape, generator=self.gen, device=self.device, dtype=torch.float64)

    def _randint(self, high, *shape):
        import torch
        return torch.randint(0, high, shape, generator=self.gen, device=self.device)

    def _gbest(self):
        """gbest position [B,D] float64: the aliased population row where the reference holds a view, else the own copy."""
        import torch
        row = torch.gather(self.pop, 1, self.gbest_row.clamp(min=0)[:, None, None].expand(self.B, 1, self.dim))[:, 0]
        return torch.where((self.gbest_row >= 0)[:, None], row, self.gbest_pos).to(torch.float64)

    @staticmethod
    def _pairwise(a, b):
        """[B,n,D], [B,m,D] float64 -> [B,n,m] Euclidean distances.  Accumulated dimension by dimension: the temporaries stay [B,n,m] (a
        broadcast difference tensor would be D times larger: 2.4 MB per lane at D = 30), and no GEMM expansion loses digits."""
        import torch
        acc = torch.zeros(a.shape[0], a.shape[1], b.shape[1], dtype=a.dtype, device=a.device)
        for d in range(a.shape[2]):
            diff = a[:, :, None, d] - b[:, None, :, d]
            acc += diff * diff
        return torch.sqrt(acc)

    def _ef(self, x, i):
        """:107-121 for particle i of every lane; x [B,NP,D] float64."""
        dist = self._pairwise(x, x).sum(dim=-1) / (NP_ - 1)
        lo, hi = dist.min(dim=1).values, dist.max(dim=1).values
        return (dist[:, i] - lo) / (hi - lo)

    def _neighbourhoods(self, alive):
        """:64-85, for the lanes still running."""
        import torch
        B, D = self.B, self.dim
        x, pb = self.pop.to(torch.float64), self.pbest_pos.to(torch.float64)
        d = self._pairwise(x, pb)                                                            # d[b, i, j] = |pbest_j - pop_i|
        d = d.masked_fill(torch.eye(NP_, dtype=torch.bool, device=self.device)[None], float('inf'))
        idx = torch.sort(d, dim=-1, stable=True).indices[:, :, :K_]
        neb = torch.gather(self.pop[:, None].expand(B, NP_, NP_, D), 2, idx[..., None].expand(B, NP_, K_, D))
        dg = torch.sqrt(torch.sum((self._gbest()[:, None, :] - x) ** 2, dim=-1))
        gidx = torch.sort(dg, dim=-1, stable=True).indices[:, :K_]
        gneb = torch.gather(self.pop, 1, gidx[..., None].expand(B, K_, D))
        m3, m4 = alive[:, None, None], alive[:, None, None, None]
        self.pbest_neb_index.copy_(torch.where(m3, idx, self.pbest_neb_index)); self.pbest_neb.copy_(torch.where(m4, neb, self.pbest_neb))
        self.gbest_neb_index.copy_(torch.where(alive[:, None], gidx, self.gbest_neb_index)); self.gbest_neb.copy_(torch.where(m3, gneb, self.gbest_neb))

    # ---- init_population (:30-61) ----------------------------------------------------------------------------------------------
    def init_population(self, tape=None):
        import torch
        B, D = self.B, self.dim
        self.last_tape = {}
        u = self._draw(tape, 'init_u', lambda: self._rand(B, NP_, D)).to(torch.float64)
        self.pop.copy_((u * (self.ub - self.lb) + self.lb).to(torch.float32))
        self.vel.zero_()
        self.cost.copy_(self._eval(self.pop, 0))
        self.pbest_pos.copy_(self.pop); self.pbest_cost.copy_(self.cost)
        self.gbest_row.copy_(torch.argmin(self.cost, dim=1))
        self.gbest_cost.copy_(self.cost.min(dim=1).values)
        self.gbest_pos.copy_(torch.gather(self.pop, 1, self.gbest_row[:, None, None].expand(B, 1, D))[:, 0])
        self.stag.zero_()
        self.r_w.copy_(self._draw(tape, 'r_w', lambda: self._rand(B)).to(torch.float64))
        self.qstate.copy_(self._draw(tape, 'qstate', lambda: self._randint(N_STATE, B, NP_)).to(torch.int64))
        self.fes.fill_(NP_); self.log_index.fill_(1); self.curve_len.fill_(1)
        self.curve.zero_(); self.curve[:, 0] = self.gbest_cost
        self.done.zero_(); self.reward.zero_()
        self.pointer, self.step_no = 0, 0
        self.feature.copy_(self.qstate[:, 0].to(torch.int32))
        return self.feature

    # ---- update (:232-297) -----------------------------------------------------------------------------------------------------
    def update(self, action, tape=None, check_action=False):
        """action [B] integer device tensor -> (feature [B] int32, reward [B] f32, done [B] uint8), all on the device, no synchronisation."""
        import torch
        B, D, i = self.B, self.dim, self.pointer
        if action.shape != (B,):
            raise ValueError(f'action error: expected shape ({B},), got {tuple(action.shape)}')
        a = action.to(device=self.device, dtype=torch.int64)
        if check_action and bool(((a < 0) | (a >= N_STATE)).any()):
            raise ValueError(f'action error: {action.tolist()}')
        a = a.clamp(0, N_STATE - 1)
        self.last_tape = {}
        self.step_no += 1
        alive = self.done == 0
        f64 = torch.float64
        if i == 0:
            self._neighbourhoods(alive)
            rw = 4 * self.r_w * (1 - self.r_w)                                                # :88-93
            frac = self.fes.to(f64) / self.maxFEs
            self.r_w.copy_(torch.where(alive, rw, self.r_w))
            self.w.copy_(torch.where(alive, U_ - (frac * rw * W_MIN + V_ * (W_MAX - W_MIN) * frac), self.w))
        # ---- generate_v_vector (:136-187) as one masked rule:  v = w v + A (P1 - x) + C (P2 - x)
        r1 = self._draw(tape, 'r1', lambda: self._rand(B)).to(f64)
        r2 = self._draw(tape, 'r2', lambda: self._rand(B)).to(f64)
        ia = self._draw(tape, 'pick_a', lambda: self._randint(K_, B)).to(torch.int64)
        ib = self._draw(tape, 'pick_b', lambda: self._randint(K_, B)).to(torch.int64)
        r1v = self._draw(tape, 'r1v', lambda: self._rand(B, D)).to(f64)
        r2v = self._draw(tape, 'r2v', lambda: self._rand(B, D)).to(f64)
        x, v = self.pop[:, i].to(f64), self.vel[:, i].to(f64)
        pb, gb = self.pbest_pos[:, i].to(f64), self._gbest()
        cs = (pb * gb).sum(-1) / (torch.sqrt((pb ** 2).sum(-1)) * torch.sqrt((gb ** 2).sum(-1)))     # :123-124
        neg = cs < 0
        p_a = torch.gather(self.gbest_neb, 1, ia[:, None, None].expand(B, 1, D))[:, 0].to(f64)
        p_b = torch.gather(self.pbest_neb[:, i], 1, ib[:, None, None].expand(B, 1, D))[:, 0].to(f64)
        c1 = torch.tensor(_C1, dtype=f64, device=self.device)[a]
        c2 = torch.tensor(_C2, dtype=f64, device=self.device)[a]
        a0, a1, a2, a3 = (a == 0), (a == 1), (a == 2), (a == 3)
        m = lambda t: t[:, None]
        P1 = torch.where(m((a0 & neg) | a2), pb, p_b)                                         # pbest for (0, cs<0) and 2; a pbest-neighbour otherwise
        P2 = torch.where(m((a0 & neg) | (a1 & ~neg) | a3), p_a, gb)                           # a gbest-neighbour for (0, cs<0), (1, cs>=0), 3; gbest otherwise
        use1 = (a0 | (a1 & neg) | (a2 & neg) | a3)                                            # strategies whose rule has the first term
        use2 = ((a0 & neg) | a1 | a2 | a3)
        k1 = torch.where(m(a3), m(c1) * r1v, m(c1 * r1).expand(B, D))
        k2 = torch.where(m(a3), m(c2) * r2v, m(c2 * r2).expand(B, D))
        zero = torch.zeros((), dtype=f64, device=self.device)
        vn = m(self.w) * v + torch.where(m(use1), k1 * (P1 - x), zero) + torch.where(m(use2), k2 * (P2 - x), zero)
        v_max = 0.1 * (self.ub - self.lb)
        vn = torch.clamp(vn, -v_max, v_max)
        am, am2 = alive[:, None], alive
        self.vel[:, i] = torch.where(am, vn.to(torch.float32), self.vel[:, i])
        xall = self.pop.to(f64)
        ef_old = self._ef(xall, i)
        x_new = torch.clamp(x + vn, self.lb, self.ub)
        xall2 = xall.clone(); xall2[:, i] = x_new
        ef_new = self._ef(xall2, i)
        x32 = x_new.to(torch.float32)
        self.pop[:, i] = torch.where(am, x32, self.pop[:, i])
        f_old = self.cost[:, i].clone()
        f_new = self._eval(self.pop[:, i:i + 1].contiguous(), 3 * self.step_no)[:, 0]
        self.fes += am2.to(torch.int32)
        cond1, cond2 = f_new < f_old, ef_new > ef_old
        rew = torch.where(cond1 & cond2, 2.0, torch.where(cond1, 1.0, torch.where(cond2, 0.0, -2.0)))      # :95-105
        self.reward.copy_(torch.where(am2, rew, torch.zeros_like(rew)).to(torch.float32))
        self.cost[:, i] = torch.where(am2, f_new, self.cost[:, i])
        better = am2 & (f_new < self.pbest_cost[:, i])                                        # (pbest_cost itself is not refreshed: reference quirk)
        self.pbest_pos[:, i] = torch.where(better[:, None], self.pop[:, i], self.pbest_pos[:, i])
        self.stag[:, i] = torch.where(better, torch.zeros_like(self.stag[:, i]), torch.where(am2, self.stag[:, i] + 1, self.stag[:, i]))
        # ---- neb_mutation (:192-230) where the particle stagnates; both mutants are evaluated for every lane, used under the mask
        mut = am2 & (self.stag[:, i] >= 2)
        mut_p = self._draw(tape, 'mut_p', lambda: self._rand(B, D)).to(f64)
        mut_g = self._draw(tape, 'mut_g', lambda: self._rand(B, D)).to(f64)
        pb = self.pbest_pos[:, i].to(f64)
        nb = self.pbest_neb[:, i].to(f64)                                                     # [B,K,D] snapshot
        order = torch.sort(torch.sqrt(torch.sum((pb[:, None, :] - nb) ** 2, dim=-1)), dim=-1, stable=True).indices
        pick = lambda t, k: torch.gather(t, 1, k[:, None, None].expand(B, 1, D))[:, 0]
        P3 = (pb + mut_p * (pick(nb, order[:, 0]) - pick(nb, order[:, -1]))).to(torch.float32)
        c3 = self._eval(P3[:, None, :].contiguous(), 3 * self.step_no + 1)[:, 0]
        self.fes += mut.to(torch.int32)
        take = mut & (c3 < self.pbest_cost[:, i])
        self.pbest_pos[:, i] = torch.where(take[:, None], P3, self.pbest_pos[:, i])
        self.pbest_cost[:, i] = torch.where(take, c3, self.pbest_cost[:, i])
        put = mut & ~take
        j = torch.gather(self.pbest_neb_index[:, i], 1, order[:, -1:])[:, 0]                  # the farthest neighbour's row is replaced
        rows = torch.arange(B, device=self.device)
        self.pop[rows, j] = torch.where(put[:, None], P3, self.pop[rows, j])
        self.cost[rows, j] = torch.where(put, c3, self.cost[rows, j])
        g = self._gbest()
        nb = self.gbest_neb.to(f64)
        order = torch.sort(torch.sqrt(torch.sum((g[:, None, :] - nb) ** 2, dim=-1)), dim=-1, stable=True).indices
        P3 = (g + mut_g * (pick(nb, order[:, 0]) - pick(nb, order[:, -1]))).to(torch.float32)
        c3 = self._eval(P3[:, None, :].contiguous(), 3 * self.step_no + 2)[:, 0]
        self.fes += mut.to(torch.int32)
        take = mut & (c3 < self.gbest_cost)
        self.gbest_pos.copy_(torch.where(take[:, None], P3, self.gbest_pos))
        self.gbest_row.copy_(torch.where(take, torch.full_like(self.gbest_row, -1), self.gbest_row))
        self.gbest_cost.copy_(torch.where(take, c3, self.gbest_cost))
        put = mut & ~take
        j = torch.gather(self.gbest_neb_index, 1, order[:, -1:])[:, 0]
        self.pop[rows, j] = torch.where(put[:, None], P3, self.pop[rows, j])
        self.cost[rows, j] = torch.where(put, c3, self.cost[rows, j])
        # ---- global best (:268-271: re-bound as a VIEW of the particle's row), strategy memory, pointer
        gb_new = am2 & (f_new < self.gbest_cost)
        self.gbest_cost.copy_(torch.where(gb_new, f_new, self.gbest_cost))
        self.gbest_row.copy_(torch.where(gb_new, torch.full_like(self.gbest_row, i), self.gbest_row))
        self.qstate[:, i] = torch.where(am2, a, self.qstate[:, i])
        self.pointer = (i + 1) % NP_
        # ---- cost-curve protocol (:276-295)
        idx = torch.arange(_lib.CURVE_LEN, device=self.device)[None, :]
        log = am2 & (self.fes >= self.log_index * self.log_interval)
        self.log_index += log.to(torch.int32)
        self.curve.copy_(torch.where(log[:, None] & (idx == self.curve_len[:, None]), self.gbest_cost[:, None], self.curve))
        self.curve_len.copy_(torch.clamp(self.curve_len + log.to(torch.int32), max=_lib.CURVE_LEN))
        fin = self.fes >= self.maxFEs
        if self.has_opt:
            fin = fin | (self.gbest_cost.to(f64) <= 1e-8)
        fin = fin & am2
        full = self.curve_len >= self.n_logpoint + 1
        pos = torch.where(full, self.curve_len - 1, self.curve_len)
        self.curve.copy_(torch.where(fin[:, None] & (idx == pos[:, None]), self.gbest_cost[:, None], self.curve))
        self.curve_len.copy_(torch.where(fin & ~full, self.curve_len + 1, self.curve_len))
        self.done.copy_(torch.where(fin, torch.ones_like(self.done), self.done))
        self.feature.copy_(torch.where(am2, self.qstate[:, self.pointer].to(torch.int32), self.feature))
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
        keys = ('pop', 'vel', 'cost', 'pbest_pos', 'pbest_cost', 'gbest_pos', 'gbest_cost', 'gbest_row', 'stag', 'qstate', 'r_w', 'w',
                'pbest_neb_index', 'pbest_neb', 'gbest_neb_index', 'gbest_neb', 'fes', 'log_index', 'curve_len', 'curve', 'done', 'feature')
        return {k: getattr(self, k)[b].cpu().numpy().copy() for k in keys}


class NRLPSO_Optimizer:
    """Drop-in for `NRLPSO_Optimizer` (nrlpso_optimizer.py:7-297): one problem, python / numpy in and out."""

    def __init__(self, config):
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
        self._eng = NRLPSO_Engine(pack, [problem._index], cfg.maxFEs, cfg.n_logpoint, self.log_interval, seed=int(np.random.randint(0, 2 ** 31 - 1)))
        self._eng.init_population()
        self._sync()
        return int(self._eng.feature[0])

    def update(self, action, problem):
        import torch
        if action not in (0, 1, 2, 3):
            raise ValueError(f'action error: {action}')
        e = self._eng
        e.update(torch.tensor([int(action)], device=e.device))
        self._sync()
        return int(e.feature[0]), int(e.reward[0]), bool(e.done[0])
