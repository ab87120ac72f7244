"""Agent layer: the rollout-time policies of the three configured MetaBBO agents, as plain torch modules
running on the same CUDA stream as the step kernels (north star: "the small src/agent policy MLP/LSTM
left as a torch call on the same stream").  Learning updates (PPO / DDQN / REINFORCE) stay out of scope.

Reference:  MLP builder `src/agent/networks.py:4-26`; RLEPSO `Actor.forward` `rlepso_agent.py:22-47`,
rollout loop `:294-303`; DE-DDQN net `de_ddqn_agent.py:21-25`, greedy action `:59-68`, rollout `:108-117`;
LDE `PolicyNet` `lde_agent.py:8-29`, rollout `:147-163`.

Module / parameter names mirror the reference so that a `state_dict` extracted from a reference pickle
(see tests/golden/make_golden_de.py: `agents.npz`) loads with `load_reference_weights`.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn


class MLP(nn.Module):
    """networks.py:4-26 -- list of {'in','out','drop_out','activation'} dicts."""

    def __init__(self, config):
        super().__init__()
        self.net = nn.Sequential()
        self.net_config = config
        for k, lc in enumerate(config):
            self.net.add_module(f'layer{k}-linear', nn.Linear(lc['in'], lc['out']))
            self.net.add_module(f'layer{k}-drop_out', nn.Dropout(lc['drop_out']))
            if lc['activation'] != 'None':
                self.net.add_module(f'layer{k}-activation', getattr(nn, lc['activation'])())

    def forward(self, x):
        return self.net(x)


def _mlp_cfg(sizes, last_act='None'):
    cfg = []
    for k in range(len(sizes) - 1):
        act = 'ReLU' if k < len(sizes) - 2 else last_act
        cfg.append({'in': sizes[k], 'out': sizes[k + 1], 'drop_out': 0, 'activation': act})
    return cfg


def load_reference_weights(module: nn.Module, weights: dict, prefix=''):
    """weights: {reference state_dict key: array}.  Name-mangled private attributes of the reference
    (`_Actor__mu_net`) are mapped onto this module's public names (`mu_net`)."""
    own = module.state_dict()
    mapped = {}
    for k, v in weights.items():
        if prefix and not k.startswith(prefix):
            continue
        kk = k[len(prefix):]
        parts = [p.split('__', 1)[1] if p.startswith('_') and '__' in p[1:] else p for p in kk.split('.')]
        kk = '.'.join(parts)
        if kk in own:
            mapped[kk] = torch.as_tensor(np.asarray(v)).to(own[kk].dtype).reshape(own[kk].shape)
    missing = set(own) - set(mapped)
    if missing:
        raise KeyError(f'reference weights missing for {sorted(missing)}')
    module.load_state_dict(mapped)
    return module


# ------------------------------------------------------------------------------------------ RLEPSO
class RLEPSO_Actor(nn.Module):
    """Two MLPs 1->64->32->35 (mu, sigma); action = clamp(Normal(mu, sigma).sample(), 0, 1)."""
    feature_dim, action_dim, max_sigma, min_sigma = 1, 35, 0.7, 0.01
    takes_noise = True          # forward(x, noise=...) accepts externally drawn standard-normal noise

    def __init__(self):
        super().__init__()
        cfg = _mlp_cfg([self.feature_dim, 64, 32, self.action_dim])
        self.mu_net = MLP(cfg)
        self.sigma_net = MLP(cfg)

    def distribution(self, x):
        mu = (torch.tanh(self.mu_net(x)) + 1.) / 2.
        sigma = (torch.tanh(self.sigma_net(x)) + 1.) / 2. * (self.max_sigma - self.min_sigma) + self.min_sigma
        return mu, sigma

    def log_prob(self, x, action):
        """sum over the 35 coefficients of log N(action; mu, sigma) (rlepso_agent.py:26-35, `fixed_action` path); differentiable."""
        mu, sigma = self.distribution(x)
        return torch.distributions.Normal(mu, sigma).log_prob(action).sum(-1)

    @torch.no_grad()
    def forward(self, x, noise=None):
        """x [..., 1] -> action [..., 35].  `noise` (standard normal, same shape as the action) can be
        injected for deterministic tests; otherwise torch.normal on the current device/stream."""
        mu, sigma = self.distribution(x)
        eps = torch.randn_like(mu) if noise is None else noise
        return torch.clamp(mu + sigma * eps, min=0, max=1)


class RLEPSO_Critic(nn.Module):
    """Value head 1 -> 16 -> 8 -> 1 (rlepso_agent.py:49-61); forward(x [..., 1]) -> value [...]."""

    def __init__(self):
        super().__init__()
        self.value_head = MLP(_mlp_cfg([RLEPSO_Actor.feature_dim, 16, 8, 1]))

    def forward(self, x):
        return self.value_head(x).squeeze(-1)


class RLEPSO_Agent:
    """Rollout side of `RLEPSO_Agent` (rlepso_agent.py:294-303) for one env or a batched engine."""

    def __init__(self, config=None, device='cuda'):
        self.device = torch.device(device)
        self.actor = RLEPSO_Actor().to(self.device).eval()

    def get_action(self, state):
        state = torch.as_tensor(np.asarray(state), dtype=torch.float32).to(self.device)
        return self.actor(state).cpu().numpy()

    def rollout_episode(self, env):
        is_done = False
        state = env.reset()
        R = 0
        while not is_done:
            action = self.get_action(state)
            state, reward, is_done = env.step(action)
            R += reward
        return {'cost': env.optimizer.cost, 'fes': env.optimizer.fes, 'return': R}


# ------------------------------------------------------------------------------------------ DE-DDQN
class DE_DDQN_Net(MLP):
    """99 -> 100 -> 100 -> 100 -> 100 -> 4 (de_ddqn_agent.py:21-25); greedy action = argmax Q."""

    def __init__(self):
        super().__init__(_mlp_cfg([99, 100, 100, 100, 100, 4]))

    @torch.no_grad()
    def act(self, state):
        return torch.argmax(self.forward(state), dim=-1)


# ------------------------------------------------------------------------------------------ LDE
class LDE_PolicyNet(nn.Module):
    """LSTM(NP+10 -> 50) + Linear mu/sigma heads -> 2*NP (lde_agent.py:8-29)."""
    takes_noise = True          # step(x, h, c, noise=...) accepts externally drawn standard-normal noise

    def __init__(self, NP=50, cell=50, bins=5):
        super().__init__()
        self.NP, self.cell = NP, cell
        self.lstm = nn.LSTM(input_size=NP + 2 * bins, hidden_size=cell, num_layers=1)
        self.mu = nn.Linear(cell, 2 * NP)
        self.sigma = nn.Linear(cell, 2 * NP)

    def forward(self, x, h, c):
        out, (h_, c_) = self.lstm(x, (h, c))
        return self.mu(out), torch.sigmoid(self.sigma(out)), h_, c_

    @torch.no_grad()
    def sampler(self, x, h, c, noise=None):
        """x [1,B,NP+10] -> action [B,2NP] clipped to [0,1], new (h,c)."""
        mu, sigma, h_, c_ = self.forward(x, h, c)
        eps = torch.randn_like(mu) if noise is None else noise
        return torch.clip(mu + sigma * eps, 0, 1)[0], h_, c_

    def lstm_cell(self, x, h, c):
        """One LSTM step written out (gate order i, f, g, o like torch.nn.LSTM) so that the rollout loop is plain
        matmuls on the current stream (CUDA-graph capturable, no cuDNN RNN descriptor).  x [B,NP+10], h/c [B,cell]
        -> mu [B,2NP], sigma [B,2NP], h', c'."""
        L = self.lstm
        gates = torch.addmm(L.bias_ih_l0 + L.bias_hh_l0, x, L.weight_ih_l0.t()) + h @ L.weight_hh_l0.t()
        i, f, g, o = gates.chunk(4, dim=-1)
        c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h2 = torch.sigmoid(o) * torch.tanh(c2)
        return self.mu(h2), torch.sigmoid(self.sigma(h2)), h2, c2

    @torch.no_grad()
    def step(self, x, h, c, noise=None):
        """Batched `sampler` for B lanes: action [B,2NP] = clip(Normal(mu, sigma).sample(), 0, 1), new (h, c)."""
        mu, sigma, h2, c2 = self.lstm_cell(x, h, c)
        eps = torch.randn_like(mu) if noise is None else noise
        return torch.clip(mu + sigma * eps, 0, 1), h2, c2


# ------------------------------------------------------------------------------------------ QLPSO
class QLPSO_QTable(nn.Module):
    """Tabular policy of QLPSO_Agent (qlpso_agent.py:17-37): a [4, 4] Q-table; the action is sampled from
    softmax(Q[state]) (`__get_action`, :33-36).  `forward(q_state [B] int, u=None) -> action [B] int32`; `u` are uniforms
    in [0, 1) (inverse-CDF sampling, so that a sweep can draw them from its own generator under CUDA-graph capture).
    `td_update` is the reference's learning rule (:47-48) for batches of transitions, applied in order."""
    takes_uniform = True

    def __init__(self, q_table=None, n_states=4, n_actions=4):
        super().__init__()
        q = torch.zeros(n_states, n_actions, dtype=torch.float64) if q_table is None else torch.as_tensor(q_table, dtype=torch.float64)
        self.register_buffer('q_table', q.clone())

    @torch.no_grad()
    def forward(self, state, u=None):
        prob = torch.softmax(self.q_table, dim=1)[state.long()]                       # exp(q) / exp(q).sum()  (:34-35)
        if u is None:
            u = torch.rand(state.shape[0], device=state.device, dtype=torch.float32)
        cdf = torch.cumsum(prob, dim=1)
        a = (u.to(cdf.dtype)[:, None] >= cdf).sum(dim=1).clamp(max=self.q_table.shape[1] - 1)
        return a.to(torch.int32)

    @torch.no_grad()
    def td_update(self, state, action, reward, next_state, alpha, gamma=0.8):
        """Q[s][a] += alpha * (r + gamma * max Q[s'] - Q[s][a]) for each transition in turn (qlpso_agent.py:47-48)."""
        q = self.q_table
        for s, a, r, n in zip(state.tolist(), action.tolist(), reward.tolist(), next_state.tolist()):
            q[s, a] += alpha * (r + gamma * q[n].max() - q[s, a])
