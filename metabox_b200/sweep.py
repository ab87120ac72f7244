"""Batched rollout drivers: the B200 counterpart of the nested `for problem / for run / while not done`
loops of `Tester.test` / `rollout` (`src/tester.py:190-219, 314-346`) and of the agents' `rollout_episode`
(`src/agent/rlepso_agent.py:294-303`, `lde_agent.py:147-163`, `de_ddqn_agent.py:108-117`).

A *lane* is one (instance, run) rollout.  Lanes are independent (fresh seed, `problem.reset()`,
`init_population` per run in the reference), so a sweep is: build the lane table -> shard it over ranks
-> every rank advances its lanes in lockstep -- per step ONE fused optimizer kernel plus the policy forward
as torch calls on the same stream, the pair captured once in a CUDA graph and replayed -- with no host
synchronisation inside the loop except a done-poll every `poll` steps -> ONE all-gather of the result rows
{curve[51], fes, return} at the end (NCCL over NVLink/NVSwitch).

Sweeps of the six backbones with a fused step kernel (graph-captured): RLEPSO, LDE, DE_DDQN, GLEET, RL_PSO, QLPSO.  Sweeps of the composed
backbones (library kernels + batched torch ops, run eagerly): DEDQN, RL_HPSDE, NRLPSO, L2L, LES.  A policy that has a single-kernel twin
(`agent.fuse`) makes the environment step two library kernels.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib
from .optimizer import DE_DDQN_Engine, GLEET_Engine, LDE_Engine, QLPSO_Engine, RL_PSO_Engine, RLEPSO_Engine


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


class _Sweep:
    """Common rollout loop.  Subclasses provide `_policy(feat) -> action` (torch, on the engine's device) and may
    keep recurrent state.  `graph=True` captures `steps_per_graph` x (policy forward + fused step + return
    accumulation) in one CUDA graph; the first call of `run()` captures, later calls only replay."""

    def __init__(self, engine, max_steps, poll=64, graph=True, steps_per_graph=1, time_kernels=False):
        self.engine = engine
        self.max_steps = int(max_steps)
        self.poll = int(poll)
        self.use_graph = bool(graph) and not time_kernels
        self.steps_per_graph = int(steps_per_graph)
        self.time_kernels = time_kernels
        self.kernel_events = []
        self.launches = 0
        self.kernels_per_step = 1           # kernels of libmbx.so per env step (2 when the policy is one of the fused kernels too)
        self.finite_reward = False          # RLEPSO: reward is +1 / -1, never NaN / inf
        self.ret = torch.zeros(engine.B, dtype=torch.float32, device=engine.device)
        self._graph = None
        # own generator for the policy's action noise: sweeps replaying their graphs concurrently on different streams
        # must not share the default generator's graph-side offset
        self.gen = torch.Generator(device=engine.device)
        self.gen.manual_seed((int(getattr(engine, 'seed', 0)) * 1000003 + int(getattr(engine, 'lane_base', 0))) & (2 ** 62 - 1))

    def _randn(self, *shape):
        return torch.randn(*shape, generator=self.gen, device=self.engine.device, dtype=torch.float32)

    # -- to be provided -------------------------------------------------------------------------------------
    def _policy(self, feat):
        raise NotImplementedError

    def _reset_policy(self):
        pass

    # -- one env step, all on the current stream, no sync -----------------------------------------------------
    def _step(self):
        eng = self.engine
        action = self._policy(eng.feature)
        if self.time_kernels:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _, reward, _ = eng.update(action)
            e1.record()
            self.kernel_events.append((e0, e1))
        else:
            _, reward, _ = eng.update(action)
        # finished lanes report reward 0 (the step kernels write it), so no alive mask is needed
        self.ret += reward if self.finite_reward else torch.nan_to_num(reward, nan=0.0, posinf=0.0, neginf=0.0)

    def _capture(self):
        # warm-up on a side stream (lazy cuBLAS / cudaFuncSetAttribute initialisation must not happen under capture)
        s = torch.cuda.Stream(device=self.engine.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._step()
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        g.register_generator_state(self.gen)
        with torch.cuda.graph(g):
            for _ in range(self.steps_per_graph):
                self._step()
        self._graph = g

    # -- the sweep in three pieces (so that several sweeps can be interleaved on different streams) -----------------
    def _begin(self):
        """init_population (+ graph capture on first use); everything on the current stream."""
        eng = self.engine
        if self.use_graph and self._graph is None:
            eng.init_population()
            self._reset_policy()
            self._capture()
        self.kernel_events = []
        eng.init_population()
        self._reset_policy()
        self.ret.zero_()
        self.launches = 1
        self.steps = 0

    def _advance(self):
        """up to `poll` more env steps on the current stream, no host sync."""
        per = self.steps_per_graph if self.use_graph else 1
        for _ in range(max(1, self.poll // per)):
            if self.steps >= self.max_steps:
                break
            if self.use_graph:
                self._graph.replay()
            else:
                self._step()
            self.launches += per * self.kernels_per_step
            self.steps += per

    def _result(self):
        eng = self.engine
        idx = torch.arange(_lib.CURVE_LEN, device=eng.device)[None, :]
        last = torch.gather(eng.curve, 1, (eng.curve_len.long() - 1).clamp(min=0)[:, None])
        curve = torch.where(idx < eng.curve_len[:, None], eng.curve, last)
        return dict(curve=curve, curve_len=eng.curve_len, fes=eng.fes, ret=self.ret, steps=self.steps)

    @torch.no_grad()
    def run(self):
        """-> dict(curve [B,51] f32 (padded like tester.py:204-205), curve_len, fes i32, ret f32) device tensors."""
        eng = self.engine
        if eng.B == 0:
            z = torch.zeros(0, device=eng.device)
            return dict(curve=torch.zeros(0, _lib.CURVE_LEN, device=eng.device), curve_len=z.int(), fes=z.int(), ret=z, steps=0)
        with torch.cuda.device(eng.device):
            self._begin()
            while self.steps < self.max_steps:
                self._advance()
                if bool(torch.all(eng.done != 0)):          # the only host sync inside the sweep
                    break
            return self._result()

    def kernel_ms(self):
        """Sum / count of the step-kernel durations recorded by the last run (needs time_kernels=True)."""
        if not self.kernel_events:
            return 0.0, 0
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in self.kernel_events), len(self.kernel_events)


class RLEPSO_Sweep(_Sweep):
    """RLEPSO_Agent.rollout_episode (rlepso_agent.py:294-303) for all lanes of one rank."""

    def __init__(self, pack, lane_inst, actor, maxFEs, n_logpoint=50, seed=0, lane_base=0, poll=64, time_kernels=False,
                 graph=True, steps_per_graph=1):
        eng = RLEPSO_Engine(pack, lane_inst, maxFEs, n_logpoint, seed=seed, lane_base=lane_base)
        # every generation charges >= NP evaluations, so ceil(maxFEs/NP) generations always suffice
        super().__init__(eng, int(math.ceil(maxFEs / eng.NP)) + 1, poll, graph, steps_per_graph, time_kernels)
        self.actor = actor
        self.finite_reward = True
        self.kernels_per_step = 2 if getattr(actor, 'fused_action', False) else 1

    def _policy(self, feat):
        if getattr(self.actor, 'takes_noise', False):
            return self.actor(feat, noise=self._randn(feat.shape[0], self.actor.action_dim))
        return self.actor(feat)


class LDE_Sweep(_Sweep):
    """LDE_Agent.rollout_episode (lde_agent.py:147-163): LSTM state (h, c) per lane, zero at reset."""

    def __init__(self, pack, lane_inst, net, maxFEs, n_logpoint=50, seed=0, lane_base=0, NP=None, poll=64,
                 time_kernels=False, graph=True, steps_per_graph=1):
        eng = LDE_Engine(pack, lane_inst, maxFEs, n_logpoint, seed=seed, lane_base=lane_base, NP=NP or net.NP)
        super().__init__(eng, int(math.ceil(maxFEs / eng.NP)) + 1, poll, graph, steps_per_graph, time_kernels)
        self.net = net
        self.h = torch.zeros(eng.B, net.cell, device=eng.device)
        self.c = torch.zeros(eng.B, net.cell, device=eng.device)

    def _reset_policy(self):
        self.h.zero_(); self.c.zero_()

    def _policy(self, feat):
        if getattr(self.net, 'takes_noise', False):
            action, h, c = self.net.step(feat, self.h, self.c, noise=self._randn(feat.shape[0], 2 * self.net.NP))
        else:
            action, h, c = self.net.step(feat, self.h, self.c)
        self.h.copy_(h); self.c.copy_(c)
        return action


class DE_DDQN_Sweep(_Sweep):
    """DE_DDQN_Agent.rollout_episode (de_ddqn_agent.py:108-117): greedy argmax of the Q-network, one FE per step.

    With the single-kernel twin of the Q-network (`agent.DE_DDQN_FusedNet`) the sweep is PERSISTENT by default: `poll` (250) environment steps
    per launch of `mbx_deddqn_run` (Q-network inside the kernel, lane state resident in shared memory; bit-identical to `poll` x (policy
    kernel, step kernel)).  `persistent=False` keeps one policy launch + one step launch per function evaluation (CUDA graph of
    `steps_per_graph` steps); a torch module as the policy always runs that way."""

    def __init__(self, pack, lane_inst, net, maxFEs, n_logpoint=50, seed=0, lane_base=0, poll=None, time_kernels=False,
                 graph=True, steps_per_graph=10, persistent=None, launches_per_poll=16):
        """`launches_per_poll` (persistent mode): launches of `poll` steps issued back to back before the host looks at `done` again.  A
        launch of ceil(B / 8) CTAs rarely fills a whole number of waves (C3: 576 CTAs on 444 resident slots); with several launches in
        flight per stream and several striped sweeps on different streams, the next launch's CTAs fill the slots the last wave leaves
        idle.  Finished lanes cost nothing (their warps skip the launch)."""
        eng = DE_DDQN_Engine(pack, lane_inst, maxFEs, n_logpoint, seed=seed, lane_base=lane_base)
        self.launches_per_poll = max(1, int(launches_per_poll))
        fused = bool(getattr(net, 'fused_action', False))
        if persistent and not fused:
            raise ValueError('persistent=True needs agent.DE_DDQN_FusedNet (the packed Q-network) as the policy')
        self.persistent = fused if persistent is None else bool(persistent)
        if poll is None:                 # persistent: 250 evaluations per launch (measured: profiles/r2r_deddqn_launch_tuning.txt); else a poll every 500 steps
            poll = 250 if self.persistent else 500
        super().__init__(eng, maxFEs - eng.NP, poll, graph and not self.persistent, steps_per_graph, time_kernels)
        self.net = net
        self.kernels_per_step = 2 if fused else 1

    def _policy(self, feat):
        if getattr(self.net, 'fused_action', False):            # agent.DE_DDQN_FusedNet: sanitising, 5 layers and argmax in one kernel
            return self.net(feat)
        q = self.net(torch.nan_to_num(feat, nan=0.0, posinf=0.0, neginf=0.0))
        return torch.argmax(q, dim=-1).to(torch.int32)

    def _advance(self):
        if not self.persistent:
            return super()._advance()
        for _ in range(self.launches_per_poll):
            n = min(self.poll, self.max_steps - self.steps)
            if n <= 0:
                return
            if self.time_kernels:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                self.engine.run_steps(self.net.packed, self.ret, n)
                e1.record()
                self.kernel_events.append((e0, e1))
            else:
                self.engine.run_steps(self.net.packed, self.ret, n)
            self.launches += 1
            self.steps += n


class DEDQN_Sweep(_Sweep):
    """DEDQN_Agent.rollout_episode (dedqn_agent.py:87-96): greedy argmax of the 4 -> 10 -> 10 -> 3 Q-network on the landscape features;
    every step charges 2 * NP evaluations, so a run is (maxFEs - 2 NP) / (2 NP) steps.  The engine is a composition of library kernels and
    torch ops with a python-side pointer, so it runs eagerly (no graph capture)."""

    def __init__(self, pack, lane_inst, net, maxFEs, n_logpoint=50, seed=0, lane_base=0, poll=16, time_kernels=False, graph=False):
        from .dedqn import DEDQN_Engine
        eng = DEDQN_Engine(pack, lane_inst, maxFEs, n_logpoint, seed=seed, lane_base=lane_base)
        super().__init__(eng, int(math.ceil(maxFEs / (2 * eng.NP))) + 1, poll, False, 1, time_kernels)
        self.net = net
        self.kernels_per_step = 3 if eng._noisy else 2    # mbx_eval (trial), mbx_eval (population; noisy suites only), mbx_landscape_features

    def _policy(self, feat):
        return torch.argmax(self.net(torch.nan_to_num(feat, nan=0.0, posinf=0.0, neginf=0.0)), dim=-1)


class NRLPSO_Sweep(_Sweep):
    """NRLPSO_Agent.rollout_episode: the strategy is sampled from softmax(Q[state]) of a 4 x 4 table (`agent.NRLPSO_QTable` or its
    single-kernel twin); one particle per step, 1 (+2 with a neighbourhood mutation) charged evaluations.  Composed engine, run eagerly."""

    def __init__(self, pack, lane_inst, policy, maxFEs, n_logpoint=50, seed=0, lane_base=0, poll=200, time_kernels=False, graph=False):
        from .nrlpso import NRLPSO_Engine
        eng = NRLPSO_Engine(pack, lane_inst, maxFEs, n_logpoint, seed=seed, lane_base=lane_base)
        super().__init__(eng, maxFEs - eng.NP, poll, False, 1, time_kernels)
        self.policy = policy
        self.finite_reward = True                    # rewards are 2 / 1 / 0 / -2
        self.kernels_per_step = 3 + (1 if getattr(policy, 'fused_action', False) else 0)     # mbx_eval: the moved particle + two mutants

    def _policy(self, feat):
        if getattr(self.policy, 'takes_uniform', False):
            u = torch.rand(feat.shape[0], generator=self.gen, device=feat.device, dtype=torch.float32)
            return self.policy(feat, u=u).reshape(-1)
        return self.policy(feat).reshape(-1)


class LES_Sweep(_Sweep):
    """LES_Agent.rollout_episode: ONE `update(action)` of the reference runs the whole episode with the agent's best network parameters
    (`params` = {'attn': [68], 'mlp': [178]}); here every generation is one engine step.  Composed engine, run eagerly (`mbx_eval`'s step
    index is a host argument)."""

    def __init__(self, pack, lane_inst, params, maxFEs, n_logpoint=50, seed=0, lane_base=0, poll=64, time_kernels=False, graph=False):
        from .les import LES_Engine, NP_
        eng = LES_Engine(pack, lane_inst, maxFEs, n_logpoint, seed=seed, lane_base=lane_base).set_params(params['attn'], params['mlp'])
        super().__init__(eng, int(math.ceil(maxFEs / NP_)) + 1, poll, False, 1, time_kernels)
        self.finite_reward = True

    def _policy(self, feat):
        return None                                   # the parameters installed at construction are the whole policy

    def _result(self):
        res = super()._result()
        res['ret'] = self.engine.episode_return().to(torch.float32)
        return res


class RL_HPSDE_Sweep(_Sweep):
    """RL_HPSDE_Agent.rollout_episode (rl_hpsde_agent.py:62-72): the action is sampled from softmax(Q[state]) of a 4 x 4 table
    (`agent.RL_HPSDE_QTable`, or its single-kernel twin); every generation charges NP + 201 evaluations with NP shrinking from 18 dim to 4.
    The engine is a composition of library kernels and torch ops with a host-side population-size schedule, so it runs eagerly."""

    def __init__(self, pack, lane_inst, policy, maxFEs, n_logpoint=50, seed=0, lane_base=0, poll=8, time_kernels=False, graph=False):
        from .rl_hpsde import NP_MIN, RW_STEPS, RL_HPSDE_Engine
        eng = RL_HPSDE_Engine(pack, lane_inst, maxFEs, n_logpoint, seed=seed, lane_base=lane_base)
        super().__init__(eng, int(math.ceil(maxFEs / (NP_MIN + RW_STEPS + 1))) + 1, poll, False, 1, time_kernels)
        self.policy = policy
        self.kernels_per_step = 4 + (1 if getattr(policy, 'fused_action', False) else 0)     # 2 x mbx_eval, walk, landscape features

    def _policy(self, feat):
        if getattr(self.policy, 'takes_uniform', False):
            u = torch.rand(feat.shape[0], generator=self.gen, device=feat.device, dtype=torch.float32)
            return self.policy(feat, u=u).reshape(-1)
        return self.policy(feat).reshape(-1)


class L2L_Sweep(_Sweep):
    """L2L_Agent.rollout_episode (l2l_agent.py:97-134): 100 steps of LSTM proposal -> evaluation; the result rows carry every second
    best-so-far value as the curve and `sum(y) / y[0]` as the return, like the reference's rollout."""

    def __init__(self, pack, lane_inst, net, maxFEs=100, n_logpoint=50, seed=0, lane_base=0, poll=50, time_kernels=False, graph=False):
        from .l2l import L2L_Engine, T_MAX
        eng = L2L_Engine(pack, lane_inst, seed=seed, lane_base=lane_base)
        super().__init__(eng, T_MAX, poll, False, 1, time_kernels)            # eager: the engine's step index is a host-side argument
        self.net = net
        self.h = torch.zeros(eng.B, net.dim, dtype=torch.float64, device=eng.device)
        self.c = torch.zeros(eng.B, net.hidden, dtype=torch.float64, device=eng.device)

    def _reset_policy(self):
        self.h.zero_(); self.c.zero_()

    def _policy(self, feat):
        h, c = self.net.step(feat, self.h, self.c)
        self.h.copy_(h); self.c.copy_(c)
        return h

    def _result(self):
        res = super()._result()
        res['ret'] = self.engine.episode_return().to(torch.float32)
        return res


class GLEET_Sweep(_Sweep):
    """GLEET_Agent.rollout_episode (gleet_agent.py:293-303): `policy(state [B,NP,27]) -> action [B,NP]` in [0,1] (the
    reference's attention actor, or any torch callable, runs on the same stream between the fused steps)."""

    def __init__(self, pack, lane_inst, policy, maxFEs, n_logpoint=50, seed=0, lane_base=0, poll=64, time_kernels=False,
                 graph=True, steps_per_graph=1):
        eng = GLEET_Engine(pack, lane_inst, maxFEs, n_logpoint, seed=seed, lane_base=lane_base)
        super().__init__(eng, int(math.ceil(maxFEs / eng.NP)) + 1, poll, graph, steps_per_graph, time_kernels)
        self.policy = policy
        self.kernels_per_step = 2 if getattr(policy, 'fused_action', False) else 1

    def _policy(self, feat):
        if getattr(self.policy, 'takes_noise', False):          # agent.GLEET_Actor: noise from the sweep's own generator (graph-safe)
            return self.policy(feat, noise=self._randn(feat.shape[0], feat.shape[1]))
        return self.policy(feat)


class RL_PSO_Sweep(_Sweep):
    """RL_PSO_Agent.rollout_episode (rl_pso_agent.py:116-124): `policy(state [B,2D]) -> action [B]`, one FE per step."""

    def __init__(self, pack, lane_inst, policy, maxFEs, n_logpoint=50, seed=0, lane_base=0, poll=500, time_kernels=False,
                 graph=True, steps_per_graph=10):
        eng = RL_PSO_Engine(pack, lane_inst, maxFEs, n_logpoint, seed=seed, lane_base=lane_base)
        super().__init__(eng, maxFEs - eng.NP, poll, graph, steps_per_graph, time_kernels)
        self.policy = policy
        self.kernels_per_step = 2 if getattr(policy, 'fused_action', False) else 1

    def _policy(self, feat):
        if getattr(self.policy, 'takes_noise', False):          # agent.RL_PSO_PolicyNet / RL_PSO_FusedPolicy
            return self.policy(feat, noise=self._randn(feat.shape[0], 1)).reshape(-1)
        return self.policy(feat).reshape(-1)


class QLPSO_Sweep(_Sweep):
    """QLPSO_Agent.rollout_episode (qlpso_agent.py:66-75): `policy(q_state [B] int32) -> action [B] integer`, one FE per step.
    A policy with `takes_uniform = True` is called as `policy(q_state, u=...)` with uniforms from the sweep's own generator
    (graph-safe sampling, see QLPSO_QTable)."""

    def __init__(self, pack, lane_inst, policy, maxFEs, n_logpoint=50, seed=0, lane_base=0, poll=500, time_kernels=False,
                 graph=True, steps_per_graph=10):
        eng = QLPSO_Engine(pack, lane_inst, maxFEs, n_logpoint, seed=seed, lane_base=lane_base)
        super().__init__(eng, maxFEs - eng.NP, poll, graph, steps_per_graph, time_kernels)
        self.policy = policy
        self.finite_reward = True                    # rewards are 2 / 1 / 0 / -2
        self.kernels_per_step = 2 if getattr(policy, 'fused_action', False) else 1

    def _policy(self, feat):
        if getattr(self.policy, 'takes_uniform', False):
            u = torch.rand(feat.shape[0], generator=self.gen, device=feat.device, dtype=torch.float32)
            return self.policy(feat, u=u).reshape(-1)
        return self.policy(feat).reshape(-1)


class StripedSweep:
    """`n_stripes` sweeps over contiguous slices of the lane table, each with its own engine, CUDA graph and CUDA stream,
    advanced round-robin.  Lanes are independent, so the stripes never synchronise with each other: the tail wave of one
    stripe's step kernel and its small policy kernels overlap with the other stripes' step kernels (the SMs stay full).
    Philox streams are keyed by the global lane id, so the results are identical to the unstriped sweep.
    `make(lane_inst_slice, lane_base) -> _Sweep`."""

    def __init__(self, make, lane_inst, lane_base=0, n_stripes=2):
        n = len(lane_inst)
        k = max(1, min(int(n_stripes), n))
        self.bounds = [shard_range(n, r, k) for r in range(k)]
        self.children = [make(lane_inst[lo:hi], lane_base + lo) for lo, hi in self.bounds]
        dev = self.children[0].engine.device
        self.streams = [torch.cuda.Stream(device=dev) for _ in self.children]
        self.engine = self.children[0].engine            # (bench.py e2e leg: lane table upload)
        self.launches = 0

    @torch.no_grad()
    def run(self):
        cur = torch.cuda.current_stream()
        for c, st in zip(self.children, self.streams):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                c._begin()
        live = list(range(len(self.children)))
        while live:
            for j in live:
                with torch.cuda.stream(self.streams[j]):
                    self.children[j]._advance()
            nxt = []
            for j in live:
                c = self.children[j]
                # the done-poll is ordered on the STRIPE's stream, behind the kernels that write `done` (on the caller's stream it would
                # race with them: a reused sweep could read last run's all-ones flags before init_population has executed)
                with torch.cuda.stream(self.streams[j]):
                    finished = bool(torch.all(c.engine.done != 0))
                if c.steps < c.max_steps and not finished:
                    nxt.append(j)
            live = nxt
        res = []
        for c, st in zip(self.children, self.streams):
            with torch.cuda.stream(st):
                r = c._result()
            for v in r.values():
                if isinstance(v, torch.Tensor):
                    v.record_stream(cur)             # allocated on the stripe's stream, consumed (torch.cat) on the caller's
            res.append(r)
            cur.wait_stream(st)
        self.launches = sum(c.launches for c in self.children)
        out = {k: torch.cat([r[k] for r in res]) for k in ('curve', 'curve_len', 'fes', 'ret')}
        out['steps'] = max(r['steps'] for r in res)
        return out


def gather_results(res, world, n_max=None):
    """The result rows [n_local, 51+2] = (curve[51], fes, return) of all ranks, in rank order, over the default process group.

    `n_max` = the largest per-rank row count, when the caller knows it (lanes sharded with `shard_range`: ceil(n / world); weak scaling:
    the common count): ONE all-gather -- the local count travels in a header row of the payload.  Without it the counts are gathered
    first (two collectives).  `fes` rides in a float32 column: exact below 2^24, and no suite charges more than 2000 * MBX_MAX_DIM =
    128 000 evaluations per run."""
    import torch.distributed as dist
    row = torch.cat([res['curve'], res['fes'].to(torch.float32)[:, None], res['ret'][:, None]], dim=1).contiguous()
    if world == 1:
        return row
    if n_max is not None:
        if row.shape[0] > n_max:
            raise ValueError(f'gather_results: {row.shape[0]} local rows > n_max = {n_max}')
        pad = torch.zeros(n_max + 1, row.shape[1], device=row.device, dtype=row.dtype)
        pad[0, 0] = float(row.shape[0])
        pad[1:1 + row.shape[0]] = row
        out = torch.empty(world * (n_max + 1), row.shape[1], device=row.device, dtype=row.dtype)
        dist.all_gather_into_tensor(out, pad)
        out = out.reshape(world, n_max + 1, row.shape[1])
        counts = [int(c) for c in out[:, 0, 0].tolist()]
        return torch.cat([out[r, 1:1 + counts[r]] for r in range(world)], dim=0)
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
