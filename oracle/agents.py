"""Oracle: rollout-time policy forward of the configured agents, torch-CPU float32 like the reference
(`src/agent/rlepso_agent.py:22-47`, `de_ddqn_agent.py:59-68`, `lde_agent.py:19-29`, `gleet_agent.py:386-440`,
`rl_pso_agent.py:22-45`).  TEST INFRASTRUCTURE ONLY.  Weights come from `tests/golden/agents.npz` / `agents2.npz`
(extracted from the reference's shipped pickles)."""
import numpy as np
import torch


def _mlp(w, prefix, x, n_layers, act=torch.relu):
    for k in range(n_layers):
        x = torch.nn.functional.linear(x, w[f'{prefix}.net.layer{k}-linear.weight'], w[f'{prefix}.net.layer{k}-linear.bias'])
        if k < n_layers - 1:
            x = act(x)
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
    return _mlp({('.' + k): v for k, v in w.items()}, '', x, 5)          # keys are 'net.layer<k>-linear.*' without a module prefix


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


# ------------------------------------------------------------------------------------------ GLEET (gleet_agent.py:386-440)
def _gleet_norm(x):
    """networks.py:68-72, normalization == 'layer': statistics over ALL [NP, E] entries of a lane, unbiased variance."""
    return (x - x.mean((1, 2)).view(-1, 1, 1)) / torch.sqrt(x.var((1, 2)).view(-1, 1, 1) + 1e-05)


def _gleet_block(w, pre, h, q=None):
    """`MultiHeadEncoder` (networks.py:256-341), head by head like `MultiHeadAttention.forward` (:156-191)."""
    q = h if q is None else q
    Wq, Wk, Wv, Wo = (w[f'{pre}.MHA_sublayer.MHA.W_{n}'] for n in ('query', 'key', 'val', 'out'))
    out = torch.zeros(q.shape[0], q.shape[1], Wo.shape[2])
    for hd in range(Wq.shape[0]):
        compat = (q @ Wq[hd]) @ (h @ Wk[hd]).transpose(1, 2) / np.sqrt(Wq.shape[2])
        out = out + (torch.softmax(compat, dim=-1) @ (h @ Wv[hd])) @ Wo[hd]
    x = _gleet_norm(out + h)
    F = f'{pre}.FFandNorm_sublayer.FF'
    ff = torch.nn.functional.linear(torch.relu(torch.nn.functional.linear(x, w[F + '.0.weight'], w[F + '.0.bias'])), w[F + '.2.weight'], w[F + '.2.bias'])
    return _gleet_norm(ff + x)


@torch.no_grad()
def gleet_distribution(w, state):
    """state float32 [B, NP, 27] -> (mu [B, NP], sigma [B, NP], decoder logits [B, NP, 16])."""
    x = torch.as_tensor(np.asarray(state), dtype=torch.float32)
    E = w['embedder.embedder.weight']
    enc = _gleet_block(w, 'encoder.0', x[:, :, :9] @ E.t())
    dec_in = torch.cat((x[:, :, 9:18] @ E.t(), x[:, :, 18:27] @ E.t()), dim=-1) @ w['embedder_for_decoder.embedder.weight'].t()
    logits = _gleet_block(w, 'decoder.0', enc, dec_in)
    lrelu = torch.nn.functional.leaky_relu
    mu = (torch.tanh(_mlp(w, 'mu_net', logits, 3, lrelu)) + 1.) / 2.
    sigma = (torch.tanh(_mlp(w, 'sigma_net', logits, 3, lrelu)) + 1.) / 2. * (0.7 - 0.01) + 0.01
    return mu[..., 0], sigma[..., 0], logits


@torch.no_grad()
def gleet_act(w, state):
    """One swarm: state [NP, 27] -> clamp(Normal(mu, sigma).sample(), 0, 1) float32 numpy [NP] (gleet_agent.py:298-300)."""
    mu, sigma, _ = gleet_distribution(w, np.asarray(state)[None])
    return torch.clamp(torch.normal(mu, sigma), min=0, max=1)[0].numpy()


# ------------------------------------------------------------------------------------------ RL_PSO (rl_pso_agent.py:22-45)
@torch.no_grad()
def rl_pso_distribution(w, state):
    x = torch.as_tensor(np.asarray(state), dtype=torch.float32)
    mu = (torch.tanh(_mlp(w, '_PolicyNetwork__mu_net', x, 3)) + 1.) / 2.
    sigma = torch.clamp((torch.tanh(_mlp(w, '_PolicyNetwork__sigma_net', x, 3)) + 1.) / 2., min=0.01, max=0.7)
    return mu, sigma


@torch.no_grad()
def rl_pso_act(w, state):
    mu, sigma = rl_pso_distribution(w, state)
    a = torch.normal(mu, sigma)
    return torch.where(torch.abs(a - 0.5) >= 0.5, (a + 3 * sigma - mu) * (1. / 6 * sigma), a).numpy()
