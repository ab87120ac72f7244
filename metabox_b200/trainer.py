"""Sharded training rollouts for the LDE agent (BASELINE config C4 "LDE train rollout ... sharded").

Reference: `LDE_Agent.train_episode` (`src/agent/lde_agent.py:85-145`): per problem, TRAJECTORY_NUM = 20 trajectories of
TRAJECTORY_LENGTH = 50 environment steps are collected with the current policy (`sampler`, no gradient), then ONE batched
forward over all stored (input, h, c) triples gives (mu, sigma), the REINFORCE loss is
`-mean(sum_k log N(action + 1e-8; mu, sigma) * discounted_reward)` with gamma = 0.99 per trajectory
(`__discounted_norm_rewards`, :68-83 -- no normalisation despite its name), one Adam step (lr 0.005), `learn_steps += 1`;
`Trainer.train` (`src/trainer.py:142-187`) loops this over the shuffled training set.

Here every (problem, trajectory) is a lane of `LDE_Engine`: the 50 steps of ALL trajectories of ALL problems of a learn
step advance in lockstep on the GPU (fused LDE step kernel + the policy's LSTM cell on the same stream), the stored
tensors never leave the device, and with torch.distributed initialised the problems of a learn step are sharded over the
ranks: local REINFORCE gradients, ONE all-reduce of the flattened gradient (~33 k floats) per learn step, identical Adam
step on every rank.  `problems_per_step = 1` on one rank reproduces the reference's update rule; a larger value averages
the per-problem losses (synchronous data parallelism).  Steps after a lane has finished (cost <= 1e-8 or maxFEs) carry
zero weight; the reference `break`s there (and its reward bookkeeping assumes that never happens within 50 steps).
"""
from __future__ import annotations

import numpy as np
import torch

from . import sweep as ms
from .agent import LDE_PolicyNet
from .optimizer import LDE_Engine

TRAJECTORY_NUM, TRAJECTORY_LENGTH, GAMMA, LR_MODEL = 20, 50, 0.99, 0.005      # lde_agent.py:36-42


def discounted_rewards(r, gamma=GAMMA):
    """r [T, L] -> discounted return-to-go per lane (lde_agent.py:68-83), on the device of r."""
    out = torch.empty_like(r)
    run = torch.zeros_like(r[0])
    for t in range(r.shape[0] - 1, -1, -1):
        run = run * gamma + r[t]
        out[t] = run
    return out


def reinforce_loss(net, inputs, actions, hs, cs, rewards, weight, lanes_per_problem=TRAJECTORY_NUM):
    """inputs [T,L,F], actions [T,L,2NP], hs/cs [T,L,C], rewards/weight [T,L]; lanes are problem-major (L = P * 20).
    -> (sum over problems of the reference's per-problem loss, number of problems).  Differentiable w.r.t. net."""
    T, L = rewards.shape
    mu, sigma, _, _ = net.lstm_cell(inputs.reshape(T * L, -1), hs.reshape(T * L, -1), cs.reshape(T * L, -1))
    logp = torch.distributions.Normal(mu, sigma).log_prob(actions.reshape(T * L, -1) + 1e-8).sum(1).reshape(T, L)   # :126-127
    disc = discounted_rewards(rewards)
    P = L // lanes_per_problem
    term = (logp * disc * weight).reshape(T, P, lanes_per_problem)
    n = weight.reshape(T, P, lanes_per_problem).sum(dim=(0, 2)).clamp(min=1.0)
    return -(term.sum(dim=(0, 2)) / n).sum(), P                                 # :129  loss = -mean(log_prob * reward)


def allreduce_gradients(params, n_local, group=None):
    """Sum the flattened gradients and the problem counts over the ranks, divide: every rank ends up with the gradient of
    the mean per-problem loss.  ONE collective per learn step.  Works on any backend (NCCL on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    params = [p for p in params if p.requires_grad]
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params] +
                     [torch.tensor([float(n_local)], device=params[0].device, dtype=params[0].dtype)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    total = float(flat[-1].item())
    off = 0
    for p in params:
        k = p.numel()
        p.grad = (flat[off:off + k] / max(total, 1.0)).reshape(p.shape).clone()
        off += k
    return total


class LDE_Trainer:
    """`LDE_Trainer(config, train_problems).train_step(problem_ids)`; config fields: dim, maxFEs, n_logpoint, log_interval."""

    def __init__(self, config, problems, net=None, NP=50, seed=0, device=None):
        from .problem import InstancePack
        self.config, self.NP, self.seed = config, int(NP), int(seed)
        self.device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.problems = list(problems)
        self.pack = InstancePack(self.problems, self.device)
        self.net = (net if net is not None else LDE_PolicyNet(NP=self.NP)).to(self.device)
        self.optimizer = torch.optim.Adam(self.net.parameters(), lr=LR_MODEL)
        self.learn_steps = 0

    @torch.no_grad()
    def collect(self, problem_ids, lane_base=0):
        """20 trajectories x 50 steps for each problem id -> dict of device tensors [T, L, ...] (+ engine for inspection)."""
        inst = np.repeat(np.asarray(problem_ids, dtype=np.int32), TRAJECTORY_NUM)
        eng = LDE_Engine(self.pack, inst, self.config.maxFEs, self.config.n_logpoint, getattr(self.config, 'log_interval', None),
                         seed=self.seed + 7919 * self.learn_steps, NP=self.NP, lane_base=lane_base)
        L, C = len(inst), self.net.cell
        feat = eng.init_population()
        h = torch.zeros(L, C, device=self.device); c = torch.zeros(L, C, device=self.device)
        I, A, H, Cc, R, W = [], [], [], [], [], []
        for _ in range(TRAJECTORY_LENGTH):
            alive = (eng.done == 0).to(torch.float32)
            x = feat.clone()
            action, h2, c2 = self.net.step(x, h, c)
            I.append(x); A.append(action); H.append(h); Cc.append(c); W.append(alive)
            feat, reward, _ = eng.update(action)
            R.append(torch.nan_to_num(reward.clone(), nan=0.0, posinf=0.0, neginf=0.0) * alive)
            h, c = h2, c2
        st = lambda xs: torch.stack(xs, dim=0)
        return dict(inputs=st(I), actions=st(A), hs=st(H), cs=st(Cc), rewards=st(R), weight=st(W), engine=eng)

    def train_step(self, problem_ids):
        """One learn step over `problem_ids` (sharded over the ranks of the default process group).
        -> {'loss', 'return', 'learn_steps', 'gbest', 'normalizer'} like train_episode's info dict (:139-142)."""
        import torch.distributed as dist
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)
        lo, hi = ms.shard_range(len(problem_ids), rank, world)
        mine = list(problem_ids[lo:hi])
        self.optimizer.zero_grad(set_to_none=True)
        info = dict(loss=0.0, ret=0.0, gbest=[], normalizer=[])
        n_local = 0
        if mine:
            tr = self.collect(mine, lane_base=lo * TRAJECTORY_NUM)
            loss, n_local = reinforce_loss(self.net, tr['inputs'], tr['actions'], tr['hs'], tr['cs'], tr['rewards'], tr['weight'])
            loss.backward()
            eng = tr['engine']
            info['loss'] = float(loss.item())
            info['ret'] = float(tr['rewards'].sum().item()) / max(1, len(mine))          # R summed over 20 trajectories (:113)
            info['gbest'] = eng.gbest.reshape(len(mine), TRAJECTORY_NUM)[:, -1].tolist()
            info['normalizer'] = eng.curve[:, 0].reshape(len(mine), TRAJECTORY_NUM)[:, -1].tolist()
        total = allreduce_gradients(self.net.parameters(), n_local)
        self.optimizer.step()
        self.learn_steps += 1
        return {'loss': info['loss'] / max(1, n_local), 'return': info['ret'], 'learn_steps': self.learn_steps,
                'gbest': info['gbest'], 'normalizer': info['normalizer'], 'problems': int(total)}
