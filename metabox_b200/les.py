"""LES backbone (a learned evolution strategy, NP = 16: self-attention over the individuals' fitness features gives the recombination
weights, a per-dimension MLP the learning rates of mean and step size) for B lanes in lockstep, and its single-problem drop-in.
Reference: `src/optimizer/les_optimizer.py:7-183`; the agent (`src/agent/les_agent.py`) meta-optimises the 68 + 178 network parameters, which
ARE the action of this backbone.  numpy twin: oracle/les.py (pinned against the unmodified reference, tests/golden/les.npz).

COMPOSED: one `mbx_eval` launch per generation ([B, 16, D], float64 in and out -- the reference keeps this tiny population in float64 and
never subtracts the optimum, and so does this engine) plus batched torch for the two small float32 networks and the float64 update rules;
nothing synchronises.  The whole episode of the reference's single `update` call is `run_episode()` here; `update()` advances one generation
(what the reference exposes through `skip_step = 1`).
"""
from __future__ import annotations

import math

import numpy as np

from . import _lib

NP_, SIGMA_RATIO = 16, 0.2                                    # les_optimizer.py:55-56
ALPHAS = (0.1, 0.5, 0.9)                                      # :52
TIMESTAMP = (1, 3, 10, 30, 50, 100, 250, 500, 750, 1000, 1250, 1500, 2000)       # :53
N_ATTN, N_MLP = 68, 178


class LES_Engine:
    """`set_params(attn [68], mlp [178])` installs the action; `update()` = one generation -> (gbest [B] f64, reward [B] (zeros: the
    reference reports one relative improvement per episode, `episode_return()`), done [B] uint8).  `tape` dicts replay draws: init `mu_u
    [B,D]`, `normal [B,16,D]`; a generation `normal [B,16,D]` (standard normals)."""
    NP = NP_

    def __init__(self, pack, lane_inst, maxFEs, n_logpoint=50, log_interval=None, seed=0, lane_base=0, _kernels=None, device=None):
        import torch
        self.pack = pack
        if _kernels is None:                                          # product path: the CUDA kernel behind the C ABI, no CPU fallback
            _lib.lib()
            self.device = pack.device
            self.dim, self.lb, self.ub = pack.dim, float(pack.problems[0].lb), float(pack.problems[0].ub)
            self.has_opt = pack.kind == _lib.KIND_BBOB
            self._eval = lambda x, gen: pack.eval_device(self.lane_inst, x, subtract_optimum=False, seed=self._eval_seed, gen=gen)
        else:                                                         # tests inject a CPU stand-in for mbx_eval
            self.device = torch.device(device or 'cpu')
            self.dim, self.lb, self.ub, self.has_opt = _kernels['dim'], float(_kernels['lb']), float(_kernels['ub']), bool(_kernels['has_opt'])
            self._eval = _kernels['eval']
        self.lane_base, self.seed = int(lane_base), int(seed)
        self._eval_seed = (self.seed * 0x9E3779B97F4A7C15 + self.lane_base * 0xD1B54A32D192ED03 + 0x1E5) & (2 ** 63 - 1)
        self.B = len(lane_inst)
        self.maxFEs, self.n_logpoint = int(maxFEs), int(n_logpoint)
        self.log_interval = int(log_interval or maxFEs // n_logpoint)
        dev, B, D = self.device, self.B, self.dim
        self.lane_inst = (lane_inst.to(device=dev, dtype=torch.int32) if isinstance(lane_inst, torch.Tensor)
                          else torch.as_tensor(np.asarray(lane_inst, dtype=np.int32)).to(dev)).contiguous()
        z = lambda *s, dt=torch.float64: torch.zeros(*s, dtype=dt, device=dev)
        self.parents, self.cost = z(B, NP_, D), z(B, NP_)
        self.mu, self.sigma, self.Pc, self.Ps = z(B, D), z(B, D), z(B, 3, D), z(B, 3, D)
        self.gbest, self.y0 = z(B), z(B)
        self.fes, self.log_index, self.curve_len = z(B, dt=torch.int32), z(B, dt=torch.int32), z(B, dt=torch.int32)
        self.curve = z(B, _lib.CURVE_LEN, dt=torch.float32)
        self.done = z(B, dt=torch.uint8)
        self.reward = z(B, dt=torch.float32)
        self.feature = self.gbest                                      # what `update` returns first (the reference returns gbest)
        self.t = 0
        self.alphas = torch.tensor(ALPHAS, dtype=torch.float64, device=dev)[None, :, None]
        self.timestamp = torch.tensor(TIMESTAMP, dtype=torch.float64, device=dev)
        self.attn = self.mlp = None
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed((self.seed * 1000003 + self.lane_base * 7919 + 53) & (2 ** 62 - 1))
        self.last_tape = {}

    def set_params(self, attn, mlp):
        """The action of the reference's `update` ({'attn': [68], 'mlp': [178]}, installed by `vector2nn` in parameter order, :7-17)."""
        import torch
        a = torch.as_tensor(np.asarray(attn, dtype=np.float32) if not isinstance(attn, torch.Tensor) else attn).to(device=self.device, dtype=torch.float32)
        m = torch.as_tensor(np.asarray(mlp, dtype=np.float32) if not isinstance(mlp, torch.Tensor) else mlp).to(device=self.device, dtype=torch.float32)
        if a.numel() != N_ATTN or m.numel() != N_MLP:
            raise AssertionError('dim of x and net not match!')                              # :8
        cut = lambda v, shapes: [t.reshape(s) for t, s in zip(torch.split(v.reshape(-1), [int(np.prod(s)) for s in shapes]), shapes)]
        self.attn = cut(a, [(8, 3), (8,), (8, 3), (8,), (1, 3), (1,)])
        self.mlp = cut(m, [(8, 19), (8,), (2, 8), (2,)])
        return self

    def _draw(self, tape, key, make):
        import torch
        if tape is not None:
            if key not in tape:
                raise KeyError(f'LES tape without {key!r}')
            a = tape[key]
            t = torch.as_tensor(np.ascontiguousarray(a) if isinstance(a, np.ndarray) else a).to(self.device)
        else:
            t = make()
        self.last_tape[key] = t
        return t

    def _sample(self, tape, mu, sigma, gen):
        import torch
        B, D = self.B, self.dim
        zn = self._draw(tape, 'normal', lambda: torch.randn(B, NP_, D, generator=self.gen, device=self.device, dtype=torch.float64)).to(torch.float64)
        pop = torch.clamp(mu[:, None, :] + sigma[:, None, :] * zn, self.lb, self.ub)
        return pop, self._eval(pop.contiguous(), gen).to(torch.float64)

    # ---- init_population (:59-82) ------------------------------------------------------------------------------------------------
    def init_population(self, tape=None):
        import torch
        B, D = self.B, self.dim
        self.last_tape = {}
        u = self._draw(tape, 'mu_u', lambda: torch.rand(B, D, generator=self.gen, device=self.device, dtype=torch.float64)).to(torch.float64)
        self.mu.copy_(self.lb + (self.ub - self.lb) * u)
        self.sigma.fill_(self.ub * SIGMA_RATIO)
        pop, cost = self._sample(tape, self.mu, self.sigma, 0)
        self.parents.copy_(pop); self.cost.copy_(cost)
        self.gbest.copy_(cost.min(dim=1).values)
        self.Pc.zero_(); self.Ps.zero_(); self.y0.zero_(); self.reward.zero_()
        self.fes.fill_(NP_); self.log_index.fill_(1); self.curve_len.fill_(1)
        self.curve.zero_(); self.curve[:, 0] = self.gbest.to(torch.float32)
        self.done.zero_()
        self.t = 0
        return None                                                     # (the reference returns None, :82)

    # ---- one pass of the loop body of update (:118-160) ----------------------------------------------------------------------------
    def update(self, action=None, tape=None):
        import torch
        if action is not None:
            self.set_params(action['attn'], action['mlp'])
        if self.attn is None:
            raise ValueError('LES_Engine.update: no network parameters installed (set_params / action)')
        B, D = self.B, self.dim
        f32, f64 = torch.float32, torch.float64
        self.last_tape = {}
        alive = self.done == 0
        lin = torch.nn.functional.linear
        c = self.cost
        zs = (c - c.mean(dim=1, keepdim=True)) / (c.std(dim=1, unbiased=False, keepdim=True) + 1e-8)            # :87-91
        rank = torch.sort(c, dim=1, stable=True).indices.to(f64) / NP_ - 0.5                                  # the permutation, not the rank
        X = torch.stack([zs, rank, (c < self.gbest[:, None]).to(f64)], dim=-1).to(f32)                         # [B,16,3]
        Wq, bq, Wk, bk, Wv, bv = self.attn
        score = torch.softmax(torch.matmul(lin(X, Wq, bq), lin(X, Wk, bk).transpose(1, 2)) / math.sqrt(8), dim=-1)
        W = torch.softmax(torch.matmul(score, lin(X, Wv, bv)), dim=1)[:, :, 0].to(f64)                          # [B,16]
        dx = self.parents - self.mu[:, None, :]
        drift = (dx * W[:, :, None]).sum(dim=1)                                                               # [B,D]
        drift_s = (dx / self.sigma[:, None, :] * W[:, :, None]).sum(dim=1)
        Pc = (1 - self.alphas) * self.Pc + self.alphas * (drift[:, None, :] - self.Pc)                        # :97-100
        Ps = (1 - self.alphas) * self.Ps + self.alphas * (drift_s[:, None, :] - self.Ps)
        rho = torch.tanh(self.t / self.timestamp - 1)[None, None, :].expand(B, D, len(TIMESTAMP))
        F = torch.cat([Pc.transpose(1, 2), Ps.transpose(1, 2), rho], dim=-1).to(f32)                           # [B,D,19]
        W1, b1, W2, b2 = self.mlp
        alpha = torch.sigmoid(lin(lin(F, W1, b1), W2, b2)).to(f64)                                            # [B,D,2]
        a_mu, a_sigma = alpha[..., 0], alpha[..., 1]
        mu = (1 - a_mu) * self.mu + a_mu * drift                                                              # :128-131
        sigma = (1 - a_sigma) * self.sigma + a_sigma * torch.sqrt((dx ** 2 * W[:, :, None]).sum(dim=1))
        self.t += 1
        pop, cost = self._sample(tape, mu, sigma, self.t)
        m1, m2, m3 = alive[:, None], alive[:, None, None], alive
        self.parents.copy_(torch.where(m2, pop, self.parents)); self.cost.copy_(torch.where(m1, cost, self.cost))
        self.mu.copy_(torch.where(m1, mu, self.mu)); self.sigma.copy_(torch.where(m1, sigma, self.sigma))
        self.Pc.copy_(torch.where(m2, Pc, self.Pc)); self.Ps.copy_(torch.where(m2, Ps, self.Ps))
        self.fes += m3.to(torch.int32) * NP_
        self.gbest.copy_(torch.where(m3, torch.minimum(cost.min(dim=1).values, self.gbest), self.gbest))
        self.y0.copy_(torch.where(m3 & (self.fes == 2 * NP_), self.gbest, self.y0))                           # init_y: gbest after the first generation
        fin = self.fes >= self.maxFEs
        if self.has_opt:
            fin = fin | (self.gbest <= 1e-8)
        fin = fin & m3
        idx = torch.arange(_lib.CURVE_LEN, device=self.device)[None, :]
        g32 = self.gbest.to(torch.float32)
        log = m3 & (self.fes >= self.log_index * self.log_interval)
        self.log_index += log.to(torch.int32)
        self.curve.copy_(torch.where(log[:, None] & (idx == self.curve_len[:, None]), g32[:, None], self.curve))
        self.curve_len.copy_(torch.clamp(self.curve_len + log.to(torch.int32), max=_lib.CURVE_LEN))
        full = self.curve_len >= self.n_logpoint + 1
        pos = torch.where(full, self.curve_len - 1, self.curve_len)
        self.curve.copy_(torch.where(fin[:, None] & (idx == pos[:, None]), g32[:, None], self.curve))
        self.curve_len.copy_(torch.where(fin & ~full, self.curve_len + 1, self.curve_len))
        self.done.copy_(torch.where(fin, torch.ones_like(self.done), self.done))
        return self.gbest, self.reward, self.done

    def episode_return(self):
        """`(init_y - gbest) / init_y` (:182)."""
        return (self.y0 - self.gbest) / self.y0

    def run_episode(self, action=None, poll=64):
        """The reference's whole `update(action)` call: generations until every lane is done (one host sync per `poll` generations)."""
        import torch
        if action is not None:
            self.set_params(action['attn'], action['mlp'])
        n_max = int(math.ceil(self.maxFEs / NP_)) + 1
        for k in range(n_max):
            self.update()
            if (k + 1) % poll == 0 and bool(torch.all(self.done != 0)):
                break
        return self.gbest, self.episode_return(), self.done

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
        keys = ('parents', 'cost', 'mu', 'sigma', 'Pc', 'Ps', 'gbest', 'fes', 'log_index', 'curve_len', 'curve', 'done')
        return {k: getattr(self, k)[b].cpu().numpy().copy() for k in keys}


class LES_Optimizer:
    """Drop-in for `LES_Optimizer` (les_optimizer.py:47-183): `init_population(problem)` -> None; `update(action, problem)` with
    action = {'attn': [68], 'mlp': [178], optional 'skip_step': n} runs the whole episode (or n generations) and returns
    (gbest, (init_y - gbest) / init_y, is_end, {})."""

    def __init__(self, config):
        self.__config = config
        self.max_fes, self.NP = config.maxFEs, NP_
        self.log_interval = config.log_interval
        self.FEs = self.cost = self.log_index = None
        self._eng = None

    def _sync(self):
        e = self._eng
        self.FEs, self.log_index = int(e.fes[0]), int(e.log_index[0])
        self.fes = self.FEs
        self.cost = [float(c) for c in e.cost_curves(pad=False)[0]]

    def init_population(self, problem):
        from .problem import InstancePack
        cfg = self.__config
        pack = problem._pack if getattr(problem, '_pack', None) is not None else InstancePack([problem])
        self._eng = LES_Engine(pack, [problem._index], cfg.maxFEs, cfg.n_logpoint, self.log_interval, seed=int(np.random.randint(0, 2 ** 31 - 1)))
        self._eng.init_population()
        self._sync()
        return None

    def update(self, action, problem):
        e = self._eng
        e.set_params(action['attn'], action['mlp'])
        skip = action.get('skip_step')
        if skip is None:
            e.run_episode()
            is_end = True
        else:
            for _ in range(int(skip)):
                e.update()
            is_end = True                                               # (the reference reports is_end = True when skip_step is reached, :149-150)
        self._sync()
        return float(e.gbest[0]), float(e.episode_return()[0]), is_end, {}
