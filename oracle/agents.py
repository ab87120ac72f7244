"""Oracle: rollout-time policy forward of the configured agents, torch-CPU float32 like the reference
(`src/agent/rlepso_agent.py:22-47`, `de_ddqn_agent.py:59-68`, `lde_agent.py:19-29`).  TEST INFRASTRUCTURE
ONLY.  Weights come from `tests/golden/agents.npz` (extracted from the reference's shipped pickles)."""
import numpy as np
import torch


def _mlp(w, prefix, x, n_layers):
    for k in range(n_layers):
        x = torch.nn.functional.linear(x, w[f'{prefix}.net.layer{k}-linear.weight'], w[f'{prefix}.net.layer{k}-linear.bias'])
        if k < n_layers - 1:
            x = torch.relu(x)
    return x


def load_weights(npz, group):
    return {k[len(group) + 1:]: torch.as_tensor(np.asarray(npz[k])) for k in npz.files if k.startswith(group + '/')}


def random_rlepso_weights(seed=0):
    g = torch.Generator().manual_seed(seed)
    w = {}
    for net in ('_Actor__mu_net', '_Actor__sigma_net'):
        for k, (i, o) in enumerate(((1, 64), (64, 32), (32, 35))):
            b = 1.0 / np.sqrt(i)
            w[f'{net}.net.layer{k}-linear.weight'] = (torch.rand(o, i, generator=g) * 2 - 1) * b
            w[f'{net}.net.layer{k}-linear.bias'] = (torch.rand(o, generator=g) * 2 - 1) * b
    return w


@torch.no_grad()
def rlepso_distribution(w, state):
    """state float32 [..,1] -> (mu, sigma) of the action distribution (rlepso_agent.py:24-25)."""
    x = torch.as_tensor(np.asarray(state), dtype=torch.float32)
    mu = (torch.tanh(_mlp(w, '_Actor__mu_net', x, 3)) + 1.) / 2.
    sigma = (torch.tanh(_mlp(w, '_Actor__sigma_net', x, 3)) + 1.) / 2. * (0.7 - 0.01) + 0.01
    return mu, sigma


@torch.no_grad()
def rlepso_act(w, state):
    """clamp(Normal(mu, sigma).sample(), 0, 1) -> float32 numpy [35]  (rlepso_agent.py:27-32, :300)."""
    mu, sigma = rlepso_distribution(w, state)
    return torch.clamp(torch.normal(mu, sigma), min=0, max=1).numpy()


@torch.no_grad()
def deddqn_q(w, state):
    x = torch.as_tensor(np.asarray(state), dtype=torch.float32)
    return _mlp(w, '', x, 5) if False else _mlp({('.' + k): v for k, v in w.items()}, '', x, 5)


@torch.no_grad()
def lde_forward(w, x, h, c):
    """LSTM(60->50) + heads (lde_agent.py:19-23): x [1,B,60], h,c [1,B,50] -> mu, sigma, h', c'."""
    lstm = torch.nn.LSTM(x.shape[-1], h.shape[-1], 1)
    lstm.weight_ih_l0.copy_(w['_PolicyNet__lstm.weight_ih_l0']); lstm.weight_hh_l0.copy_(w['_PolicyNet__lstm.weight_hh_l0'])
    lstm.bias_ih_l0.copy_(w['_PolicyNet__lstm.bias_ih_l0']); lstm.bias_hh_l0.copy_(w['_PolicyNet__lstm.bias_hh_l0'])
    out, (h_, c_) = lstm(x, (h, c))
    mu = torch.nn.functional.linear(out, w['_PolicyNet__mu.weight'], w['_PolicyNet__mu.bias'])
    sigma = torch.sigmoid(torch.nn.functional.linear(out, w['_PolicyNet__sigma.weight'], w['_PolicyNet__sigma.bias']))
    return mu, sigma, h_, c_
