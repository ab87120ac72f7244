"""Sharded training rollouts for the LDE agent (REINFORCE) and the RLEPSO / GLEET agents (PPO) (BASELINE config C4
"LDE train rollout ... sharded").  The PPO part is at the end of the file.

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
from .agent import LDE_PolicyNet, RLEPSO_Actor, RLEPSO_Critic
from .optimizer import GLEET_Engine, LDE_Engine, RLEPSO_Engine

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


def allreduce_gradients(params, n_local, group=None, extra=None):
    """Sum the flattened gradients and the problem counts over the ranks, divide: every rank ends up with the gradient of
    the mean per-problem loss.  ONE collective per learn step.  Works on any backend (NCCL on GPUs, gloo in the CPU tests).
    `extra`: optional list of local scalars summed in the same collective -> returns (total, [sums])."""
    import torch.distributed as dist
    params = [p for p in params if p.requires_grad]
    tail = [float(v) for v in (extra or [])] + [float(n_local)]
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params] +
                     [torch.tensor(tail, device=params[0].device, dtype=params[0].dtype)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    sums = flat[-len(tail):].tolist()
    total = float(sums[-1])
    off = 0
    for p in params:
        k = p.numel()
        p.grad = (flat[off:off + k] / max(total, 1.0)).reshape(p.shape).clone()
        off += k
    return total if extra is None else (total, sums[:-1])


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
    def collect(self, problem_ids, lane_base=0, time_kernels=False):
        """20 trajectories x 50 steps for each problem id -> dict of device tensors [T, L, ...] (+ engine for inspection).
        `time_kernels`: CUDA events around every fused-step launch (`kernel_events` in the result; measurement only)."""
        inst = np.repeat(np.asarray(problem_ids, dtype=np.int32), TRAJECTORY_NUM)
        eng = LDE_Engine(self.pack, inst, self.config.maxFEs, self.config.n_logpoint, getattr(self.config, 'log_interval', None),
                         seed=self.seed + 7919 * self.learn_steps, NP=self.NP, lane_base=lane_base)
        L, C = len(inst), self.net.cell
        feat = eng.init_population()
        h = torch.zeros(L, C, device=self.device); c = torch.zeros(L, C, device=self.device)
        I, A, H, Cc, R, W, events = [], [], [], [], [], [], []
        for _ in range(TRAJECTORY_LENGTH):
            alive = (eng.done == 0).to(torch.float32)
            x = feat.clone()
            action, h2, c2 = self.net.step(x, h, c)
            I.append(x); A.append(action); H.append(h); Cc.append(c); W.append(alive)
            if time_kernels:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                feat, reward, _ = eng.update(action)
                e1.record()
                events.append((e0, e1))
            else:
                feat, reward, _ = eng.update(action)
            R.append(torch.nan_to_num(reward.clone(), nan=0.0, posinf=0.0, neginf=0.0) * alive)
            h, c = h2, c2
        st = lambda xs: torch.stack(xs, dim=0)
        return dict(inputs=st(I), actions=st(A), hs=st(H), cs=st(Cc), rewards=st(R), weight=st(W), engine=eng, kernel_events=events)

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


# ======================================================================================== RLEPSO (PPO)
N_STEP, K_EPOCHS, EPS_CLIP, PPO_GAMMA, PPO_LR = 10, 3, 0.1, 0.999, 1e-5          # rlepso_agent.py:72-79


def ppo_targets(rewards, weight, bootstrap, gamma=PPO_GAMMA):
    """n-step targets (rlepso_agent.py:219-231): R_t = r_t + gamma R_{t+1}, R_T = critic(state after the segment), per lane.
    rewards / weight [T, L], bootstrap [L] (detached).  Steps after a lane has finished (weight 0) leave its chain untouched."""
    out = torch.empty_like(rewards)
    run = bootstrap
    for t in range(rewards.shape[0] - 1, -1, -1):
        run = torch.where(weight[t] > 0, run * gamma + rewards[t], run)
        out[t] = run
    return out


def ppo_losses(actor, critic, states, actions, old_logp, rewards, weight, next_state, old_value=None, eps_clip=EPS_CLIP,
               gamma=PPO_GAMMA):
    """One K-epoch of the reference's PPO update (rlepso_agent.py:196-253) for L lanes at once.
    states [T,L,1], actions [T,L,35], old_logp / rewards / weight [T,L], next_state [L,1]; old_value: detached values of the
    first epoch (None in the first epoch).  -> (reinforce_loss, baseline_loss, values.detach(), n_lanes): each loss is the SUM
    over the lanes with at least one live step of the reference's per-episode mean over that lane's steps."""
    logp = actor.log_prob(states, actions)                                   # [T, L]
    value = critic(states)                                                   # [T, L]
    with torch.no_grad():
        target = ppo_targets(rewards, weight, critic(next_state), gamma)
    ratios = torch.exp(logp - old_logp.detach())
    adv = target - value.detach()
    surr = torch.min(ratios * adv, torch.clamp(ratios, 1 - eps_clip, 1 + eps_clip) * adv)
    if old_value is None:
        verr = (value - target) ** 2
    else:
        vclip = old_value + torch.clamp(value - old_value, -eps_clip, eps_clip)
        verr = torch.max((value - target) ** 2, (vclip - target) ** 2)
    n_t = weight.sum(0)
    live = n_t > 0
    n_t = n_t.clamp(min=1.0)
    reinforce = -((surr * weight).sum(0) / n_t)[live].sum()
    baseline = ((verr * weight).sum(0) / n_t)[live].sum()
    return reinforce, baseline, value.detach(), int(live.sum().item())


class _PPOTrainer:
    """The segment loop shared by the PPO agents (RLEPSO, GLEET -- the reference's two `train_episode`s are the same algorithm):
    every problem is a lane of a batched engine, all lanes advance in lockstep in segments of `n_step` generations (policy
    sampling + fused step kernel on the GPU, nothing leaves the device), and after each segment `k_epochs` clipped-PPO updates
    are made from the mean of the per-lane losses.  With torch.distributed initialised the problems are sharded over the ranks
    and ALL gradients travel in ONE all-reduce per update (the same collective carries the number of lanes still running, so
    every rank leaves the loop in the same segment).  One problem on one rank is the reference's update rule.
    Subclasses provide: `_engine`, `_state`, `_sample`, `_heads`, `_params`, `_zero_grad`, `_backward`, `_apply`."""
    n_step, k_epochs, eps_clip, gamma = N_STEP, K_EPOCHS, EPS_CLIP, PPO_GAMMA

    def _setup(self, config, problems, seed, device):
        from .problem import InstancePack
        self.config, self.seed = config, int(seed)
        self.device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.problems = list(problems)
        self.pack = InstancePack(self.problems, self.device)
        self.learning_time = 0
        self.episodes = 0

    @torch.no_grad()
    def collect_segment(self, eng, n_step=None):
        """n_step generations of all lanes with the current policy -> states [T,L,...], actions [T,L,...], logp, rewards, weight [T,L]."""
        S, A, LP, R, W = [], [], [], [], []
        for _ in range(n_step or self.n_step):
            alive = (eng.done == 0).to(torch.float32)
            x = self._state(eng)
            a, lp = self._sample(x)
            S.append(x); A.append(a); LP.append(lp); W.append(alive)
            _, reward, _ = eng.update(a)
            R.append(torch.nan_to_num(reward.clone().to(torch.float32), nan=0.0, posinf=0.0, neginf=0.0) * alive)
        st = lambda xs: torch.stack(xs, dim=0)
        return dict(states=st(S), actions=st(A), logp=st(LP), rewards=st(R), weight=st(W), next_state=self._state(eng))

    def train_episodes(self, problem_ids):
        """Full episodes of `problem_ids` (sharded over the ranks).  -> info dict like train_episode's with per-problem lists
        plus 'segments' and the last losses."""
        import torch.distributed as dist
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)
        lo, hi = ms.shard_range(len(problem_ids), rank, world)
        mine = np.asarray(list(problem_ids[lo:hi]), dtype=np.int32)
        max_ls = getattr(self.config, 'max_learning_step', None)
        eng = self._engine(mine, lo)
        if len(mine):
            eng.init_population()
        ret = torch.zeros(len(mine), device=self.device)
        actor, critic = self._heads()
        segments, last, stop = 0, (0.0, 0.0), False
        running = 1.0
        while running > 0 and not stop:
            seg = self.collect_segment(eng) if len(mine) else None
            if seg is not None:
                ret += seg['rewards'].sum(0)
            old_value = None
            for _k in range(self.k_epochs):
                self._zero_grad()
                n_local, alive_after = 0, 0.0
                if seg is not None:
                    rl, bl, val, n_local = ppo_losses(actor, critic, seg['states'], seg['actions'], seg['logp'], seg['rewards'],
                                                      seg['weight'], seg['next_state'], old_value, self.eps_clip, self.gamma)
                    if old_value is None:
                        old_value = val
                    if n_local:
                        self._backward(rl, bl)
                        last = (float(rl.item()) / n_local, float(bl.item()) / n_local)
                    alive_after = float((eng.done == 0).sum().item())
                total, (running,) = allreduce_gradients(self._params(), n_local, extra=[alive_after])
                if total > 0:
                    self._apply()
                    self.learning_time += 1
                if max_ls is not None and self.learning_time >= max_ls:
                    stop = True
                    break
            segments += 1
        self.episodes += 1
        curves = eng.cost_curves(pad=False) if len(mine) else []
        return {'normalizer': [c[0] for c in curves], 'gbest': [c[-1] for c in curves], 'return': ret.tolist(),
                'learn_steps': self.learning_time, 'segments': segments, 'reinforce_loss': last[0], 'baseline_loss': last[1],
                'done': bool(max_ls is not None and self.learning_time >= max_ls)}


class RLEPSO_Trainer(_PPOTrainer):
    """`RLEPSO_Trainer(config, train_problems).train_episodes(problem_ids)` -- `RLEPSO_Agent.train_episode`
    (rlepso_agent.py:113-292) for many problems at once (see `_PPOTrainer`): segments of n_step = 10 generations, K_epochs = 3
    clipped-PPO updates of actor and critic (two Adam optimizers, lr 1e-5, two separate backward passes, :257-258).
    config fields: dim, maxFEs, n_logpoint, log_interval, optional max_learning_step."""

    def __init__(self, config, problems, actor=None, critic=None, seed=0, device=None):
        self._setup(config, problems, seed, device)
        self.actor = (actor if actor is not None else RLEPSO_Actor()).to(self.device)
        self.critic = (critic if critic is not None else RLEPSO_Critic()).to(self.device)
        self.optimizer_actor = torch.optim.Adam(self.actor.parameters(), lr=PPO_LR)
        self.optimizer_critic = torch.optim.Adam(self.critic.parameters(), lr=PPO_LR)

    def _params(self):
        return list(self.actor.parameters()) + list(self.critic.parameters())

    def _heads(self):
        return self.actor, self.critic

    def _engine(self, mine, lo):
        cfg = self.config
        return RLEPSO_Engine(self.pack, mine, cfg.maxFEs, cfg.n_logpoint, getattr(cfg, 'log_interval', None),
                             seed=self.seed + 104729 * self.episodes, lane_base=lo)

    def _state(self, eng):
        return eng.feature.reshape(eng.B, 1).clone()

    def _sample(self, x):
        a = self.actor(x)
        return a, self.actor.log_prob(x, a)

    def _zero_grad(self):
        self.optimizer_actor.zero_grad(set_to_none=True)
        self.optimizer_critic.zero_grad(set_to_none=True)

    def _backward(self, rl, bl):
        bl.backward()
        rl.backward()

    def _apply(self):
        self.optimizer_actor.step()
        self.optimizer_critic.step()


# ------------------------------------------------------------------------------------------------------------- GLEET (PPO)
GLEET_LR, GLEET_MAX_GRAD_NORM = 1e-4, 0.1                                            # gleet_agent.py:50-52


class _GleetActorHead:
    """`ppo_losses` view of the GLEET actor: states [T,L,NP,27], actions [T,L,NP] -> joint log-probability [T,L].  The decoder output of
    the last `log_prob` call is kept (with its autograd graph) so that the critic head, which `ppo_losses` calls next on the SAME states,
    does not run the attention stack a second time."""

    def __init__(self, actor):
        self.actor = actor
        self.cached = (None, None)                      # (states tensor, decoder logits [T*L, NP, 16])

    def log_prob(self, states, actions):
        lead = states.shape[:-2]
        mu, sigma, lg = self.actor.distribution(states.reshape(-1, *states.shape[-2:]))
        self.cached = (states, lg)
        return torch.distributions.Normal(mu, sigma).log_prob(actions.reshape(-1, actions.shape[-1])).sum(-1).reshape(lead)


class _GleetCriticHead:
    """The GLEET critic reads the ACTOR's decoder output (gleet_agent.py:147-150, 236): value(states [...,NP,27]) -> [...]; the baseline loss
    therefore also trains the actor's attention stack, exactly like the reference."""

    def __init__(self, actor, critic, actor_head=None):
        self.actor, self.critic, self.actor_head = actor, critic, actor_head

    def __call__(self, states):
        lead = states.shape[:-2]
        if self.actor_head is not None and self.actor_head.cached[0] is states:
            lg = self.actor_head.cached[1]
            self.actor_head.cached = (None, None)       # one use: the next epoch runs with updated weights
        else:
            lg = self.actor.logits(states.reshape(-1, *states.shape[-2:]))
        return self.critic(lg).reshape(lead)


class GLEET_Trainer(_PPOTrainer):
    """`GLEET_Agent.train_episode` (gleet_agent.py:111-290) for many problems at once (see `_PPOTrainer`): segments of n_step = 10
    generations, K_epochs = 3 clipped-PPO epochs (eps 0.1, gamma 0.999, value clipping from the second epoch on), ONE Adam optimizer over
    actor + critic (lr 1e-4; the reference builds an ExponentialLR scheduler but never steps it), loss = baseline + reinforce in one
    backward pass, gradient-norm clipping at 0.1 per parameter group (actor, critic) after the all-reduce (`clip_grad_norms`, :472-487)."""

    def __init__(self, config, problems, actor=None, critic=None, seed=0, device=None, fused_collection=False):
        """`fused_collection=True`: trajectories are sampled by the single-kernel actor (`agent.GLEET_FusedActor`: action, mu, sigma in one
        launch per generation, log-probability from those) instead of the torch module; the packed weights are refreshed after every
        optimiser step.  The update epochs always differentiate through the torch module."""
        from .agent import GLEET_Actor, GLEET_Critic
        self._setup(config, problems, seed, device)
        self.actor = (actor if actor is not None else GLEET_Actor()).to(self.device)
        self.critic = (critic if critic is not None else GLEET_Critic()).to(self.device)
        self.optimizer = torch.optim.Adam([{'params': self.actor.parameters(), 'lr': GLEET_LR}, {'params': self.critic.parameters(), 'lr': GLEET_LR}])
        self._fused = None
        if fused_collection:
            from .agent import GLEET_FusedActor
            self._fused = GLEET_FusedActor(self.actor)

    def _params(self):
        return list(self.actor.parameters()) + list(self.critic.parameters())

    def _heads(self):
        head = _GleetActorHead(self.actor)
        return head, _GleetCriticHead(self.actor, self.critic, head)

    def _engine(self, mine, lo):
        cfg = self.config
        return GLEET_Engine(self.pack, mine, cfg.maxFEs, cfg.n_logpoint, getattr(cfg, 'log_interval', None),
                            seed=self.seed + 104729 * self.episodes, lane_base=lo)

    def _state(self, eng):
        return eng.feature.clone()

    def _sample(self, x):
        if self._fused is not None:
            a, mu, sigma, _ = self._fused(x, noise=torch.randn(x.shape[:2], device=x.device), outputs=True)
            return a, torch.distributions.Normal(mu, sigma).log_prob(a).sum(-1)
        a = self.actor(x)
        return a, self.actor.log_prob(x, a)

    def _zero_grad(self):
        self.optimizer.zero_grad(set_to_none=True)

    def _backward(self, rl, bl):
        (bl + rl).backward()

    def _apply(self):
        norms = [torch.nn.utils.clip_grad_norm_(group['params'], GLEET_MAX_GRAD_NORM, norm_type=2) for group in self.optimizer.param_groups]
        # GLEET's state divides costs by `max_cost` = the MINIMUM of the initial costs (gleet_optimizer.py:49, reproduced), so features reach
        # 1e4 on badly scaled functions at dim 30 and a batch of such lanes can overflow the backward pass.  One non-finite gradient would
        # turn every weight into NaN (clip_grad_norm_ scales by NaN): such an update is skipped and counted instead.
        if not all(bool(torch.isfinite(n)) for n in norms):
            self.skipped_updates = getattr(self, 'skipped_updates', 0) + 1
            self.optimizer.zero_grad(set_to_none=True)
            return
        self.optimizer.step()
        if self._fused is not None:
            self._fused.refresh()
